// dann_group.h — query-batch data parallelism inside ONE host process: a group owns one full
// replica of the index per GPU; a batch is cut into contiguous slices (one per device), every device runs the whole
// hot path on its slice, and the "gather" is each device's D2H copy landing directly in the caller's arrays at the
// slice's offset - rows come back in query order without any device-to-device exchange (queries are independent
// units; the graph itself is replicas only, sharding it would put every hop on NVLink).  This is what a Rust host
// (one Postgres backend, or the sidecar that serves many) can call; torch.distributed / NCCL are only needed when
// ranks are separate PROCESSES (pgvectorscale_b200/group.py, bench.py under torchrun).
// Included at the end of diskann_b200.cu (it uses search_batch_host, fail() and DANN_CATCH); pure host code.
#pragma once
#include <condition_variable>
#include <thread>

struct GroupWorker {
    dann_index *ix = nullptr;
    std::thread th;
    /* the slice of the current call */
    int lo = 0, hi = 0;
    int rc = 0;
    std::string err;
    uint64_t seen = 0; /* generation this worker has finished */
};

struct dann_group {
    std::vector<GroupWorker> w;
    std::mutex mu, call_mu; /* call_mu: one batch call at a time per group */
    std::condition_variable cv_go, cv_done;
    uint64_t gen = 0;
    bool stop = false;
    /* arguments of the current call (borrowed from the caller for its duration) */
    const float *queries = nullptr;
    const int16_t *labels = nullptr;
    const int32_t *label_off = nullptr;
    int k = 0, L = 0, rescore = 0;
    uint32_t dim = 0;
    uint64_t *out_tid = nullptr;
    float *out_dist = nullptr;
    uint32_t *out_count = nullptr;
    dann_query_stats *out_stats = nullptr;

    void run(size_t i) {
        GroupWorker &me = w[i];
        uint64_t done = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_go.wait(lk, [&] { return stop || gen != done; });
            if (stop) return;
            const uint64_t g = gen;
            const int lo = me.lo, hi = me.hi;
            lk.unlock();
            int rc = DANN_OK;
            std::string err;
            if (hi > lo) {
                try {
                    rc = search_batch_host(me.ix, queries + (size_t)lo * dim, labels, label_off ? label_off + lo : nullptr, hi - lo,
                                           k, L, rescore, out_tid + (size_t)lo * k, out_dist ? out_dist + (size_t)lo * k : nullptr,
                                           nullptr, out_count ? out_count + lo : nullptr, out_stats ? out_stats + lo : nullptr);
                    if (rc) err = g_err; /* thread-local message of this worker */
                } catch (const std::bad_alloc &) {
                    rc = DANN_ERR_OOM;
                } catch (...) {
                    rc = DANN_ERR_STATE;
                }
            }
            lk.lock();
            me.rc = rc;
            try {
                me.err = err;
            } catch (...) {
            }
            done = g;
            me.seen = g;
            cv_done.notify_all();
        }
    }
};

/* contiguous, balanced slice [lo, hi) of `total` queries for device `r` of `world` (the first ranks take the remainder;
 * the same rule as pgvectorscale_b200/group.py shard_bounds) */
static void group_bounds(int total, int world, int r, int *lo, int *hi) {
    const int base = total / world, rem = total % world;
    *lo = r * base + std::min(r, rem);
    *hi = *lo + base + (r < rem ? 1 : 0);
}

extern "C" void dann_group_free(dann_group *g) {
    if (!g) return;
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->stop = true;
    }
    g->cv_go.notify_all();
    for (GroupWorker &x : g->w)
        if (x.th.joinable()) x.th.join();
    for (GroupWorker &x : g->w) dann_index_free(x.ix);
    delete g;
}

extern "C" int dann_group_create(const dann_snapshot_desc *snap, int ndev, const int *devices, dann_group **out) try {
    if (!snap || !out || ndev < 1 || ndev > 64) return fail(DANN_ERR_INVALID_ARG, "dann_group_create: bad argument");
    dann_group *g = new dann_group();
    g->w.resize((size_t)ndev);
    for (int i = 0; i < ndev; i++) {
        const int dev = devices ? devices[i] : i;
        int rc = dann_index_load(snap, dev, &g->w[(size_t)i].ix);
        if (rc) {
            const std::string keep = g_err;
            dann_group_free(g); /* frees the replicas loaded so far */
            return fail(rc, "dann_group_create: replica %d (device %d): %s", i, dev, keep.c_str());
        }
    }
    g->dim = g->w[0].ix->v.dim;
    for (size_t i = 0; i < g->w.size(); i++) g->w[i].th = std::thread([g, i] { g->run(i); });
    *out = g;
    return DANN_OK;
} DANN_CATCH

extern "C" int dann_group_size(const dann_group *g) { return g ? (int)g->w.size() : 0; }

extern "C" dann_index *dann_group_replica(dann_group *g, int i) {
    return (g && i >= 0 && (size_t)i < g->w.size()) ? g->w[(size_t)i].ix : nullptr;
}

extern "C" int dann_group_search_batch(dann_group *g, const float *queries, const int16_t *labels, const int32_t *label_off,
                                       int B, int k, int search_list_size, int rescore, uint64_t *out_tid, float *out_dist,
                                       uint32_t *out_count, dann_query_stats *out_stats) try {
    if (!g || B <= 0 || k <= 0 || !queries || !out_tid) return fail(DANN_ERR_INVALID_ARG, "dann_group_search_batch: bad argument");
    std::lock_guard<std::mutex> one(g->call_mu);
    std::unique_lock<std::mutex> lk(g->mu);
    g->queries = queries;
    g->labels = labels;
    g->label_off = label_off;
    g->k = k;
    g->L = search_list_size;
    g->rescore = rescore;
    g->out_tid = out_tid;
    g->out_dist = out_dist;
    g->out_count = out_count;
    g->out_stats = out_stats;
    const int world = (int)g->w.size();
    for (int r = 0; r < world; r++) group_bounds(B, world, r, &g->w[(size_t)r].lo, &g->w[(size_t)r].hi);
    const uint64_t gen = ++g->gen;
    g->cv_go.notify_all();
    g->cv_done.wait(lk, [&] {
        for (const GroupWorker &x : g->w)
            if (x.seen != gen) return false;
        return true;
    });
    for (int r = 0; r < world; r++)
        if (g->w[(size_t)r].rc) return fail(g->w[(size_t)r].rc, "replica %d: %s", r, g->w[(size_t)r].err.c_str());
    return DANN_OK;
} DANN_CATCH
