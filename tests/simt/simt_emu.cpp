// simt_emu.cpp — fibers, scheduler and warp collectives of the SIMT emulator (see simt_emu.h).  Test infrastructure.
#include <execinfo.h>
#include <mutex>
#include "simt_emu.h"

#include <algorithm>

#if !defined(__x86_64__)
#error "the fiber switch below is written for x86-64 System V"
#endif

extern "C" void simt_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch, .-simt_switch
)");

namespace simt {

Block *g_blk = nullptr;
Fiber *g_cur = nullptr;
static uint64_t g_total_switches = 0;
static uint64_t g_coll_parity[2] = {0, 0}; /* lane-participations in completed warp collectives (32 = one full-warp collective)
                                              by warp index parity (two-warp kernel: 0 = controller, 1 = heap) */
static const size_t kStack = 128 * 1024;

uint64_t total_switches() { return g_total_switches; }
void collectives_by_warp_parity(uint64_t out[2]) {
    out[0] = g_coll_parity[0] / 32;
    out[1] = g_coll_parity[1] / 32;
}

void yield() {
    Block *b = g_blk;
    Fiber *f = g_cur;
    b->switches++;
    g_total_switches++;
    simt_switch(&f->sp, b->sched_sp);
}

static void complete(Slot &s, unsigned nlanes_mask);

static void trampoline() {
    Fiber *f = g_cur;
    g_blk->body();
    f->done = true;
    g_blk->progress++;
    { /* a thread that has returned no longer takes part in its warp's collectives */
        Warp &w = g_blk->warps[f->tid >> 5];
        w.exited |= 1u << (f->tid & 31u);
        for (Slot &c : w.slots)
            if (c.arrived && ((c.arrived | w.exited) & c.mask) == c.mask) {
                complete(c, c.mask);
                for (int l = 0; l < 32; l++)
                    if ((c.arrived >> l) & 1u) {
                        w.result[l] = c.out[c.gen & 1][l];
                        w.ready[l] = true;
                    }
                c.arrived = 0;
                c.gen++;
                g_blk->progress++;
                g_blk->collectives++;
            }
    }
    yield();
    fprintf(stderr, "simt: resumed a finished fiber\n");
    abort();
}

static void prepare(Fiber &f, char *stack) {
    f.stack = stack;
    uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
    void **sp = (void **)top;
    *--sp = nullptr;             /* where a caller's return address would sit: keeps rsp = 8 (mod 16) at entry */
    *--sp = (void *)trampoline;  /* `ret` target of the first switch */
    for (int i = 0; i < 6; i++) *--sp = nullptr; /* rbp rbx r12 r13 r14 r15 */
    f.sp = sp;
}

static void complete(Slot &s, unsigned nlanes_mask) {
    uint64_t *out = s.out[s.gen & 1];
    const unsigned mask = s.mask;
    (void)nlanes_mask;
    switch (s.kind) {
        case K_SYNC:
            break;
        case K_BALLOT: {
            unsigned r = 0;
            for (int l = 0; l < 32; l++)
                if (((s.arrived >> l) & 1u) && s.val[l]) r |= 1u << l;
            for (int l = 0; l < 32; l++) out[l] = r;
            break;
        }
        case K_ANY:
        case K_ALL: {
            bool any = false, all = true;
            for (int l = 0; l < 32; l++)
                if ((s.arrived >> l) & 1u) {
                    any |= s.val[l] != 0;
                    all &= s.val[l] != 0;
                }
            for (int l = 0; l < 32; l++) out[l] = s.kind == K_ANY ? any : all;
            break;
        }
        case K_SHFL:
        case K_SHFL_UP:
        case K_SHFL_DOWN:
        case K_SHFL_XOR:
            for (int l = 0; l < 32; l++) {
                if (!((s.arrived >> l) & 1u)) continue;
                int src = l;
                if (s.kind == K_SHFL) src = (int)(s.aux[l] & 31u);
                else if (s.kind == K_SHFL_UP) src = l - (int)s.aux[l] >= 0 ? l - (int)s.aux[l] : l;
                else if (s.kind == K_SHFL_DOWN) src = l + (int)s.aux[l] <= 31 ? l + (int)s.aux[l] : l;
                else src = (int)((unsigned)l ^ s.aux[l]) & 31;
                if (!((mask >> src) & 1u)) {
                    fprintf(stderr, "simt: shuffle reads lane %d which is not in mask %08x\n", src, mask);
                    abort();
                }
                /* reading a lane that has exited is undefined in CUDA; give the reader its own value back */
                out[l] = ((s.arrived >> src) & 1u) ? s.val[src] : s.val[l];
            }
            break;
        case K_REDUX_MIN:
        case K_REDUX_ADD: {
            uint64_t r = s.kind == K_REDUX_MIN ? ~0ull : 0ull;
            for (int l = 0; l < 32; l++)
                if ((s.arrived >> l) & 1u) r = s.kind == K_REDUX_MIN ? (s.val[l] < r ? s.val[l] : r) : (uint64_t)(uint32_t)(r + s.val[l]);
            for (int l = 0; l < 32; l++) out[l] = r;
            break;
        }
        case K_MATCH_ANY:
            for (int l = 0; l < 32; l++) {
                if (!((s.arrived >> l) & 1u)) continue;
                unsigned r = 0;
                for (int m = 0; m < 32; m++)
                    if (((s.arrived >> m) & 1u) && s.val[m] == s.val[l]) r |= 1u << m;
                out[l] = r;
            }
            break;
        default:
            abort();
    }
}

uint64_t rendezvous(int kind, unsigned mask, uint64_t val, uint32_t aux) {
    Block *b = g_blk;
    Fiber *f = g_cur;
    const unsigned lane = f->tid & 31u;
    Warp &w = b->warps[f->tid >> 5];
    if (!((mask >> lane) & 1u)) {
        fprintf(stderr, "simt: lane %u calls a collective with mask %08x that excludes it\n", lane, mask);
        abort();
    }
    Slot *s = nullptr;
    for (Slot &c : w.slots)
        if (c.arrived && c.mask == mask && !((c.arrived >> lane) & 1u)) {
            s = &c;
            break;
        }
    if (!s)
        for (Slot &c : w.slots)
            if (!c.arrived) {
                s = &c;
                s->mask = mask;
                s->kind = kind;
                break;
            }
    if (!s) {
        fprintf(stderr, "simt: more than 32 divergent collectives in flight in one warp\n");
        abort();
    }
    if (s->kind != kind) {
        fprintf(stderr, "simt: lanes of warp %u meet in different collectives (%d vs %d, mask %08x)\n", f->tid >> 5, s->kind,
                kind, mask);
        {   /* where this lane is (resolve with addr2line -e <the emulator .so> -f -C -i <offsets>) */
            void *bt[24];
            const int nbt = backtrace(bt, 24);
            backtrace_symbols_fd(bt, nbt, 2);
        }
        abort();
    }
    const uint64_t mygen = s->gen;
    s->val[lane] = val;
    s->aux[lane] = aux;
    s->arrived |= 1u << lane;
    if (((s->arrived | w.exited) & mask) == mask) {
        complete(*s, mask);
        for (int l = 0; l < 32; l++)
            if ((s->arrived >> l) & 1u) {
                w.result[l] = s->out[mygen & 1][l];
                w.ready[l] = true;
            }
        s->arrived = 0;
        s->gen++;
        b->progress++;
        b->collectives++;
        g_coll_parity[(f->tid >> 5) & 1u] += (uint64_t)__builtin_popcount(mask);
    } else {
        f->blocked_on = "warp collective";
        while (!w.ready[lane]) yield(); /* the slot itself may already serve another group's collective by now */
        f->blocked_on = nullptr;
    }
    w.ready[lane] = false;
    return w.result[lane];
}

void named_barrier(unsigned id, unsigned count) {
    Block *b = g_blk;
    Fiber *f = g_cur;
    if (id >= 16) abort();
    Barrier &br = b->bars[id];
    if (br.arrived == 0) br.count = count;
    else if (br.count != count) {
        fprintf(stderr, "simt: bar.sync %u with different thread counts (%u vs %u)\n", id, br.count, count);
        abort();
    }
    const uint64_t mygen = br.gen;
    br.arrived++;
    if (br.arrived == count) {
        br.arrived = 0;
        br.gen++;
        b->progress++;
    } else {
        f->blocked_on = "named barrier";
        while (br.gen == mygen) yield();
        f->blocked_on = nullptr;
    }
}

void launch(unsigned grid, unsigned block, const std::function<void()> &body) {
    /* the emulator's state (current block, fiber switcher) is global: launches from several host threads - the
     * replicas of a dann_group, one worker thread each - run one after the other */
    static std::mutex launch_mu;
    std::lock_guard<std::mutex> launch_lock(launch_mu);
    /* one set of fiber stacks per launch, reused by every block (blocks run one after the other) */
    char *stacks = (char *)aligned_alloc(64, (size_t)block * kStack);
    if (!stacks) abort();
    for (unsigned bi = 0; bi < grid; bi++) {
        Block b;
        b.bidx.x = bi;
        b.bdim.x = block;
        b.gdim.x = grid;
        b.body = body;
        b.fibers.resize(block);
        b.warps.resize((block + 31) / 32);
        for (unsigned t = 0; t < block; t++) {
            b.fibers[t].tid = t;
            b.fibers[t].tidx.x = t;
            prepare(b.fibers[t], stacks + (size_t)t * kStack);
        }
        g_blk = &b;
        unsigned live = block;
        uint64_t last_progress = ~0ull;
        int idle_passes = 0;
        /* fiber visiting order per pass: SIMT_SCHED=0 ascending thread ids, 1 descending, 2 a fresh pseudo-random
         * permutation every pass (seed SIMT_SEED).  Collectives make results independent of the order unless the
         * kernel has an intra-warp race (a shared-memory exchange without __syncwarp), which one of the orders
         * then exposes. */
        const char *se = getenv("SIMT_SCHED");
        const int sched = se ? atoi(se) : 0;
        const char *sd = getenv("SIMT_SEED");
        uint64_t rng = (sd ? strtoull(sd, nullptr, 10) : 1) * 0x9E3779B97F4A7C15ull + bi;
        std::vector<unsigned> order(block);
        for (unsigned t = 0; t < block; t++) order[t] = sched == 1 ? block - 1 - t : t;
        while (live) {
            if (sched == 2)
                for (unsigned t = block; t > 1; t--) {
                    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                    std::swap(order[t - 1], order[(rng >> 33) % t]);
                }
            for (unsigned oi = 0; oi < block; oi++) {
                const unsigned t = order[oi];
                Fiber &f = b.fibers[t];
                if (f.done) continue;
                g_cur = &f;
                simt_switch(&b.sched_sp, f.sp);
                if (f.done) live--;
            }
            if (b.progress == last_progress) {
                if (++idle_passes > 4) {
                    fprintf(stderr, "simt: deadlock in block %u:", bi);
                    for (unsigned t = 0; t < block; t++)
                        if (!b.fibers[t].done) fprintf(stderr, " t%u(%s)", t, b.fibers[t].blocked_on ? b.fibers[t].blocked_on : "running");
                    fprintf(stderr, "\n");
                    abort();
                }
            } else {
                idle_passes = 0;
                last_progress = b.progress;
            }
        }
        g_blk = nullptr;
        g_cur = nullptr;
    }
    free(stacks);
}

}  // namespace simt
