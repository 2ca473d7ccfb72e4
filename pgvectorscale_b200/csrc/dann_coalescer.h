// dann_coalescer.h — query coalescing for process-per-connection hosts.
//
// Postgres runs one backend per connection and `amcanparallel = false` (mod.rs:63): every backend has exactly one
// index scan in flight, one query at a time, while the GPU reaches its throughput on batches (one query = one warp
// pair).  The coalescer is the piece a sidecar that owns the HBM-resident index would put between the two: any number
// of host threads (one per connected backend) submit single queries and block; a dispatcher thread gathers what
// arrives within a short window into ONE dann_search_batch call and hands every caller its own rows.  Queries are
// independent units, so each caller gets exactly what a private scan would have returned.
// Included at the end of diskann_b200.cu (it uses search_batch_host and fail()); pure host code.
#pragma once
#include <chrono>
#include <condition_variable>
#include <deque>
#include <thread>

struct CoalescedRequest {
    const float *query;
    const int16_t *labels;
    int nlabels; /* < 0: no scan key */
    int k, L, rescore;
    uint64_t *out_tid;
    float *out_dist;
    uint32_t *out_count;
    dann_query_stats *out_stats;
    int rc = 0;
    bool done = false;
    std::string err;
};

struct dann_coalescer {
    dann_index *ix = nullptr;
    int max_batch = 256;
    int max_wait_us = 200;
    std::mutex mu;
    std::condition_variable cv_work, cv_done, cv_idle;
    int callers = 0; /* threads inside dann_coalescer_search (guarded by mu) */
    std::deque<CoalescedRequest *> queue;
    bool stop = false;
    std::thread worker;
    uint64_t n_batches = 0, n_queries = 0, max_seen = 0;

    static bool compatible(const CoalescedRequest *a, const CoalescedRequest *b) {
        /* one dann_search_batch call shares k / L / rescore and is either all keyed or all unkeyed */
        return a->k == b->k && a->L == b->L && a->rescore == b->rescore && (a->nlabels >= 0) == (b->nlabels >= 0);
    }

    void run() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_work.wait(lk, [&] { return stop || !queue.empty(); });
            if (stop && queue.empty()) return;
            /* the window opens with the first request: wait for company, but never longer than max_wait_us */
            const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(max_wait_us);
            cv_work.wait_until(lk, deadline, [&] { return stop || (int)queue.size() >= max_batch; });
            std::vector<CoalescedRequest *> batch;
            CoalescedRequest *head = queue.front();
            try {
                batch.reserve(std::min<size_t>(queue.size(), (size_t)max_batch));
            } catch (...) { /* no memory for the batch list: fail the oldest request, keep serving */
                queue.pop_front();
                head->rc = DANN_ERR_OOM;
                head->err = "host allocation failed";
                head->done = true;
                cv_done.notify_all();
                continue;
            }
            for (auto it = queue.begin(); it != queue.end() && (int)batch.size() < max_batch;) {
                if (compatible(head, *it)) {
                    batch.push_back(*it); /* reserved above: does not allocate */
                    it = queue.erase(it);
                } else {
                    ++it; /* different scan parameters: next window */
                }
            }
            lk.unlock();
            /* the dispatcher is a std::thread: an exception escaping it would std::terminate the host process (the
             * sidecar and every backend behind it), so whatever execute() throws becomes the batch's status */
            int xrc = 0;
            try {
                execute(batch);
            } catch (const std::bad_alloc &) {
                xrc = DANN_ERR_OOM;
            } catch (...) {
                xrc = DANN_ERR_STATE;
            }
            lk.lock();
            n_batches++;
            n_queries += batch.size();
            max_seen = std::max<uint64_t>(max_seen, batch.size());
            for (CoalescedRequest *r : batch) {
                if (xrc) {
                    r->rc = xrc;
                    try {
                        r->err = xrc == DANN_ERR_OOM ? "host allocation failed" : "unexpected C++ exception in the coalescer";
                    } catch (...) {
                    }
                }
                r->done = true;
            }
            cv_done.notify_all();
        }
    }

    void execute(std::vector<CoalescedRequest *> &batch) {
        const int B = (int)batch.size();
        const uint32_t dim = ix->v.dim;
        const int k = batch[0]->k;
        const bool keyed = batch[0]->nlabels >= 0;
        std::vector<float> q((size_t)B * dim);
        std::vector<int16_t> lab;
        std::vector<int32_t> off;
        for (int b = 0; b < B; b++) memcpy(q.data() + (size_t)b * dim, batch[b]->query, (size_t)dim * sizeof(float));
        if (keyed) {
            off.assign(1, 0);
            for (int b = 0; b < B; b++) {
                lab.insert(lab.end(), batch[b]->labels, batch[b]->labels + batch[b]->nlabels);
                off.push_back((int32_t)lab.size());
            }
            if (lab.empty()) lab.push_back(0);
        }
        std::vector<uint64_t> tid((size_t)B * k);
        std::vector<float> dist((size_t)B * k);
        std::vector<uint32_t> count(B);
        std::vector<dann_query_stats> stats(B);
        const int rc = search_batch_host(ix, q.data(), keyed ? lab.data() : nullptr, keyed ? off.data() : nullptr, B, k,
                                         batch[0]->L, batch[0]->rescore, tid.data(), dist.data(), nullptr, count.data(),
                                         stats.data());
        const std::string err = rc ? g_err : std::string();
        for (int b = 0; b < B; b++) {
            CoalescedRequest *r = batch[b];
            r->rc = rc;
            r->err = err;
            if (rc) continue;
            memcpy(r->out_tid, tid.data() + (size_t)b * k, (size_t)k * sizeof(uint64_t));
            if (r->out_dist) memcpy(r->out_dist, dist.data() + (size_t)b * k, (size_t)k * sizeof(float));
            if (r->out_count) *r->out_count = count[b];
            if (r->out_stats) *r->out_stats = stats[b];
        }
    }
};

extern "C" int dann_coalescer_create(dann_index *ix, int max_batch, int max_wait_us, dann_coalescer **out) try {
    if (!ix || !out) return fail(DANN_ERR_INVALID_ARG, "dann_coalescer_create: NULL argument");
    if (max_batch < 1 || max_batch > 65536 || max_wait_us < 0) return fail(DANN_ERR_INVALID_ARG, "dann_coalescer_create: bad limits");
    dann_coalescer *c = new (std::nothrow) dann_coalescer();
    if (!c) return fail(DANN_ERR_OOM, "host allocation failed");
    c->ix = ix;
    c->max_batch = max_batch;
    c->max_wait_us = max_wait_us;
    c->worker = std::thread([c] { c->run(); });
    *out = c;
    return DANN_OK;
} DANN_CATCH

extern "C" int dann_coalescer_search(dann_coalescer *c, const float *query, const int16_t *labels, int nlabels, int k,
                                     int search_list_size, int rescore, uint64_t *out_tid, float *out_dist,
                                     uint32_t *out_count, dann_query_stats *out_stats) try {
    if (!c || !query || !out_tid || k <= 0) return fail(DANN_ERR_INVALID_ARG, "dann_coalescer_search: bad argument");
    if (nlabels > 0 && !labels) return fail(DANN_ERR_INVALID_ARG, "labels is NULL but nlabels > 0");
    CoalescedRequest r;
    r.query = query;
    r.labels = labels;
    r.nlabels = nlabels;
    r.k = k;
    r.L = search_list_size;
    r.rescore = rescore;
    r.out_tid = out_tid;
    r.out_dist = out_dist;
    r.out_count = out_count;
    r.out_stats = out_stats;
    std::unique_lock<std::mutex> lk(c->mu);
    if (c->stop) return fail(DANN_ERR_STATE, "coalescer is shutting down");
    /* `callers` keeps dann_coalescer_destroy from deleting the object while this thread still holds (or is about to
     * re-acquire, inside cv_done.wait) its mutex */
    struct Leave {
        dann_coalescer *c;
        ~Leave() {
            if (--c->callers == 0) c->cv_idle.notify_all();
        }
    };
    c->callers++;
    Leave leave{c}; /* destroyed before lk: runs under the lock */
    c->queue.push_back(&r);
    c->cv_work.notify_one();
    c->cv_done.wait(lk, [&] { return r.done; });
    if (r.rc) return fail(r.rc, "%s", r.err.c_str());
    return DANN_OK;
} DANN_CATCH

extern "C" int dann_coalescer_stats(dann_coalescer *c, uint64_t *batches, uint64_t *queries, uint64_t *largest_batch) try {
    if (!c) return fail(DANN_ERR_INVALID_ARG, "NULL coalescer");
    std::lock_guard<std::mutex> lk(c->mu);
    if (batches) *batches = c->n_batches;
    if (queries) *queries = c->n_queries;
    if (largest_batch) *largest_batch = c->max_seen;
    return DANN_OK;
} DANN_CATCH

extern "C" void dann_coalescer_destroy(dann_coalescer *c) {
    if (!c) return;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        c->stop = true;
    }
    c->cv_work.notify_all();
    if (c->worker.joinable()) c->worker.join(); /* drains the queue: every queued request is answered */
    {
        std::unique_lock<std::mutex> lk(c->mu);
        c->cv_idle.wait(lk, [&] { return c->callers == 0; }); /* the answered callers have let go of the mutex */
    }
    delete c;
}
