// simt_emu.h — a small SIMT emulator: runs the repo's CUDA kernel SOURCES (pgvectorscale_b200/csrc/*.cuh) on the
// CPU, one fiber per CUDA thread, so that kernel LOGIC (warp collectives, named barriers, the two-warp hand-off,
// the Rust-heap emulation) can be checked against the oracle on a box without a GPU.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under pgvectorscale_b200/ includes, links or loads this; the product has no
// CPU path (tests/test_abi.py::test_no_cpu_fallback_without_device).  What it is for: catching logic errors
// before GPU minutes are spent, and keeping alternative kernel code paths that cannot be timed yet bit-exact.
//
// Model.  A block is a set of fibers (hand-switched x86-64 contexts, one OS thread).  A fiber runs until it
// reaches a warp collective / barrier whose other participants have not arrived, then yields; collectives
// complete when every lane of the mask has arrived.  Lanes therefore run maximally OUT of lockstep between
// collectives (legal under independent thread scheduling), which flags code that silently relies on
// warp-synchronous execution.  Not modelled: memory ordering weaker than sequential consistency, timing.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace simt {

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
};

enum Kind { K_SYNC = 1, K_BALLOT, K_SHFL, K_SHFL_UP, K_SHFL_DOWN, K_SHFL_XOR, K_MATCH_ANY, K_ANY, K_ALL, K_REDUX_MIN, K_REDUX_ADD };

struct Slot {
    unsigned mask = 0, arrived = 0;
    uint64_t gen = 0;
    int kind = 0;
    uint64_t val[32];
    uint32_t aux[32];
    uint64_t out[2][32];
};

struct Warp {
    Slot slots[32]; /* collectives in flight at once (lane groups with disjoint masks meet independently) */
    uint64_t result[32]; /* per-lane mailbox: a lane is in at most one collective, so its result waits here safely */
    bool ready[32] = {};
    unsigned exited = 0; /* lanes whose thread has returned: collectives do not wait for them (CUDA: "all NON-EXITED
                            threads named in mask must execute the same intrinsic") */
};

struct Barrier {
    unsigned arrived = 0, count = 0;
    uint64_t gen = 0;
};

struct Fiber {
    void *sp = nullptr;
    char *stack = nullptr;
    bool done = false, started = false;
    unsigned tid = 0;
    Dim3 tidx;
    const char *blocked_on = nullptr;
};

struct Block {
    Dim3 bidx, bdim, gdim;
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    Barrier bars[16];
    std::function<void()> body;
    void *sched_sp = nullptr;
    uint64_t progress = 0;
    uint64_t switches = 0, collectives = 0;
};

extern Block *g_blk;
extern Fiber *g_cur;

void yield();
uint64_t rendezvous(int kind, unsigned mask, uint64_t val, uint32_t aux);
void named_barrier(unsigned id, unsigned count);
/* runs `body` once per thread of every block of the grid, blocks one after the other */
void launch(unsigned grid, unsigned block, const std::function<void()> &body);
/* the same with the launch's dynamic shared memory size checked against the emulator's arena (dann_smem, 256 KB) */
inline void launch_checked(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body) {
    if (smem_bytes > 256u * 1024u) {
        fprintf(stderr, "simt: launch asks for %zu bytes of dynamic shared memory, the arena holds 262144\n", smem_bytes);
        abort();
    }
    launch(grid, block, body);
}
uint64_t total_switches();
void collectives_by_warp_parity(uint64_t out[2]);

inline unsigned lane_id() { return g_cur->tid & 31u; }

template <typename T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload too wide");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace simt

/* ------------------------------------------------------------------------------------------------
 * CUDA surface used by the kernels in pgvectorscale_b200/csrc (search path)
 * ---------------------------------------------------------------------------------------------- */
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __shared__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__

#define threadIdx (::simt::g_cur->tidx)
#define blockIdx (::simt::g_blk->bidx)
#define blockDim (::simt::g_blk->bdim)
#define gridDim (::simt::g_blk->gdim)

struct __attribute__((aligned(8))) uint2 { /* CUDA's alignment: a misaligned pair load traps on the GPU */
    uint32_t x, y;
};
struct __attribute__((aligned(16))) uint4 {
    uint32_t x, y, z, w;
};
struct __attribute__((aligned(16))) ulonglong2 {
    unsigned long long x, y;
};
struct __attribute__((aligned(16))) float4 {
    float x, y, z, w;
};
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

template <typename T>
inline T min(T a, T b) {
    return b < a ? b : a;
}
template <typename T>
inline T max(T a, T b) {
    return a < b ? b : a;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline int __float_as_int(float f) { return ::simt::from_bits<int>(::simt::to_bits(f)); }
inline float __int_as_float(int i) { return ::simt::from_bits<float>(::simt::to_bits(i)); }
template <typename T>
inline T __ldg(const T *p) {
    return *p;
}
template <typename T>
inline T __ldcg(const T *p) {
    return *p;
}
/* explicit-rounding f32 arithmetic: plain IEEE ops (the emulator is built with -ffp-contract=off) */
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
inline float __fsqrt_rn(float a) { return __builtin_sqrtf(a); } /* IEEE sqrt is correctly rounded */
inline float __ull2float_rn(unsigned long long v) { return (float)v; }
inline void __threadfence_block() {}
inline void __threadfence() {}

inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { ::simt::rendezvous(::simt::K_SYNC, mask, 0, 0); }
inline void __syncthreads() {
    ::simt::named_barrier(0, ::simt::g_blk->bdim.x * ::simt::g_blk->bdim.y * ::simt::g_blk->bdim.z);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    return (unsigned)::simt::rendezvous(::simt::K_BALLOT, mask, pred ? 1u : 0u, 0);
}
inline int __any_sync(unsigned mask, int pred) { return (int)::simt::rendezvous(::simt::K_ANY, mask, pred ? 1u : 0u, 0); }
inline int __all_sync(unsigned mask, int pred) { return (int)::simt::rendezvous(::simt::K_ALL, mask, pred ? 1u : 0u, 0); }
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    if (width != 32) abort();
    return ::simt::from_bits<T>(::simt::rendezvous(::simt::K_SHFL, mask, ::simt::to_bits(v), (uint32_t)src));
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    if (width != 32) abort();
    return ::simt::from_bits<T>(::simt::rendezvous(::simt::K_SHFL_UP, mask, ::simt::to_bits(v), delta));
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    if (width != 32) abort();
    return ::simt::from_bits<T>(::simt::rendezvous(::simt::K_SHFL_DOWN, mask, ::simt::to_bits(v), delta));
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    if (width != 32) abort();
    return ::simt::from_bits<T>(::simt::rendezvous(::simt::K_SHFL_XOR, mask, ::simt::to_bits(v), (uint32_t)lanemask));
}
inline unsigned __reduce_min_sync(unsigned mask, unsigned v) {
    return (unsigned)::simt::rendezvous(::simt::K_REDUX_MIN, mask, v, 0);
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    return (unsigned)::simt::rendezvous(::simt::K_REDUX_ADD, mask, v, 0);
}
template <typename T>
inline unsigned __match_any_sync(unsigned mask, T v) {
    return (unsigned)::simt::rendezvous(::simt::K_MATCH_ANY, mask, ::simt::to_bits(v), 0);
}

/* one OS thread: plain read-modify-write is atomic here */
template <typename T>
inline T atomicAdd(T *p, T v) {
    T o = *p;
    *p = o + v;
    return o;
}
template <typename T>
inline T atomicSub(T *p, T v) {
    T o = *p;
    *p = o - v;
    return o;
}
template <typename T>
inline T atomicOr(T *p, T v) {
    T o = *p;
    *p = o | v;
    return o;
}
template <typename T>
inline T atomicAnd(T *p, T v) {
    T o = *p;
    *p = o & v;
    return o;
}
template <typename T>
inline T atomicCAS(T *p, T cmp, T v) {
    T o = *p;
    if (o == cmp) *p = v;
    return o;
}
template <typename T>
inline T atomicMax(T *p, T v) {
    T o = *p;
    if (v > o) *p = v;
    return o;
}
template <typename T>
inline T atomicExch(T *p, T v) {
    T o = *p;
    *p = v;
    return o;
}
