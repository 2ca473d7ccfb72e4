// dann_heap.cuh — device clone of Rust's std::collections::BinaryHeap sift rules.
//
// The reference keeps its candidates in `BinaryHeap<Reverse<ListSearchNeighbor>>`
// (graph/mod.rs:75,144-147,166) and its rerank window in `BinaryHeap<ResortData>` with a
// reversed Ord (scan.rs:91-117,279-304).  SBQ Hamming distances are small integers, so ties
// are everywhere and the ORDER IN WHICH TIES POP is whatever std's heap does; returned row
// ids are only bit-exact if the device runs the same algorithm
// (alloc::collections::binary_heap, Rust 1.7x-1.8x):
//
//   push  = Vec::push + sift_up(0, old_len)
//   pop   = Vec::pop, swap with data[0], sift_down_to_bottom(0)
//   sift_up(start,pos): move the hole up while !(elem <= parent)
//   sift_down_to_bottom(pos): walk the hole to the bottom taking
//         child += (data[child] <= data[child+1])      (right child on ties)
//       then sift_up(start, hole)
//
// Both reference heaps are "min on key": for Reverse<T> and for ResortData,
// `a <= b` <=> key(b) <= key(a).  An entry packs the key in its high bits (KSHIFT) and a
// payload below; only the key takes part in comparisons (ties compare Equal:
// neighbor_with_distance.rs:31-43,74-83 for candidates, scan.rs:111-117 for the window).
#pragma once
#include <stdint.h>

template <typename E, int KSHIFT>
struct RustHeap {
    static __device__ __forceinline__ uint32_t key(E e) { return (uint32_t)(e >> KSHIFT); }

    template <typename Store>
    static __device__ __forceinline__ uint32_t sift_up(Store &s, uint32_t start, uint32_t pos,
                                                       E elem) {
        const uint32_t k = key(elem);
        while (pos > start) {
            uint32_t parent = (pos - 1) >> 1;
            E pe = s.get(parent);
            if (key(pe) <= k) break; /* hole.element() <= hole.get(parent) */
            s.set(pos, pe);
            pos = parent;
        }
        s.set(pos, elem);
        return pos;
    }

    template <typename Store>
    static __device__ __forceinline__ void push(Store &s, uint32_t &len, E elem) {
        sift_up(s, 0, len, elem);
        len++;
    }

    /* ---- warp-cooperative forms (all 32 lanes call them with identical arguments) --------
     * sift_up(0, pos) walks ONE root-ward path: lane j fetches the ancestor at height j+1
     * (1-based index (pos+1) >> (j+1)), a ballot finds the first ancestor with key <= elem
     * (where std's loop breaks), every ancestor below it moves one level down and the
     * element lands in the freed slot.  Same final array as the sequential loop, at the cost
     * of one load + one ballot + one store whatever the rise height. */
    template <typename Store>
    static __device__ __forceinline__ void sift_up_warp(Store &s, uint32_t pos, E elem, int lane) {
        const uint32_t k = key(elem);
        const uint32_t p1 = pos + 1;
        const uint32_t a1 = lane < 31 ? (p1 >> (lane + 1)) : 0u; /* 1-based ancestor, 0 = none */
        E av = 0;
        if (a1) av = s.get(a1 - 1);
        const unsigned above = __ballot_sync(0xFFFFFFFFu, a1 != 0 && key(av) > k);
        const int rise = __ffs(~above) - 1; /* consecutive ancestors (from the parent up) the element passes */
        /* lanes below `rise` move their ancestor one level down, lane `rise` drops the element */
        if (lane <= rise) s.set((p1 >> lane) - 1, lane < rise ? av : elem);
        __syncwarp();
    }

    /* pop(): lane 0 walks the hole to the bottom (data-dependent path), then the displaced
     * last element is sifted up cooperatively.  len must be > 0; returns nothing because the
     * caller already knows the root (it peeked it). */
    template <typename Store>
    static __device__ __forceinline__ void pop_warp(Store &s, uint32_t &len, int lane) {
        len--;
        if (len == 0) return;
        uint32_t pos = 0;
        E item = 0;
        if (lane == 0) {
            item = s.get(len);
            const uint32_t end = len;
            const uint32_t lim = end >= 2 ? end - 2 : 0;
            uint32_t child = 1;
            while (child <= lim) {
                E cl = s.get(child), cr = s.get(child + 1);
                if (key(cr) <= key(cl)) { /* data[child] <= data[child+1]: right child on ties */
                    child++;
                    cl = cr;
                }
                s.set(pos, cl);
                pos = child;
                child = 2 * pos + 1;
            }
            if (child == end - 1) {
                s.set(pos, s.get(child));
                pos = child;
            }
        }
        pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        __syncwarp();
        sift_up_warp(s, pos, item, lane);
    }

    /* ---- 1-based storage forms (slot p holds Rust's data[p-1]; parent p>>1, children 2p, 2p+1).
     * Child pairs are then 8-byte aligned, so the pop's descent reads both children with one
     * 64-bit load, and the ancestor index of lane j is simply p >> j. ------------------------ */
    /* PSM: the caller guarantees that every PARENT slot on the path is in shared memory (true
     * whenever the leaf slot is < 2*hs), so only lane 0's leaf store may go to the HBM tail. */
    template <bool PSM = false, typename Store>
    static __device__ __forceinline__ void sift_up_warp1(Store &s, uint32_t p, E elem, int lane) {
        const uint32_t k = key(elem);
        const uint32_t ast = p >> lane; /* slot at height `lane` on the path (0 = above the root) */
        const uint32_t ald = ast >> 1;  /* its parent */
        E av = 0;
        if (ald) av = PSM ? s.get_sm(ald) : s.get(ald);
        const unsigned above = __ballot_sync(0xFFFFFFFFu, ald != 0 && key(av) > k);
        const unsigned tm = above & ~(above + 1u); /* trailing ones: the ancestors the element passes */
        const unsigned wm = tm | (tm + 1u);        /* lanes 0..rise write */
        if ((wm >> lane) & 1u) {
            const E val = ((tm >> lane) & 1u) ? av : elem;
            if (PSM && lane > 0) s.set_sm(ast, val);
            else s.set(ast, val);
        }
        __syncwarp();
    }

    /* len = number of elements (slots 1..len), must be > 0 */
    template <typename Store>
    static __device__ __forceinline__ void pop_warp1(Store &s, uint32_t &len, int lane) {
        len--;
        if (len == 0) return;
        uint32_t p = 1;
        E item = 0;
        if (lane == 0) {
            item = s.get(len + 1);
            const uint32_t end = len;
            uint32_t c = 2;
            while (c + 1 <= end) { /* Rust: child <= end.saturating_sub(2) */
                E cl, cr;
                s.get2(c, cl, cr);
                if (key(cr) <= key(cl)) { /* data[child] <= data[child+1]: right child on ties */
                    c++;
                    cl = cr;
                }
                s.set(p, cl);
                p = c;
                c = 2 * p;
            }
            if (c == end) { /* child == end - 1 */
                s.set(p, s.get(c));
                p = c;
            }
        }
        p = __shfl_sync(0xFFFFFFFFu, p, 0);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        __syncwarp();
        sift_up_warp1<false>(s, p, item, lane);
    }

    /* len must be > 0 */
    template <typename Store>
    static __device__ __forceinline__ E pop(Store &s, uint32_t &len) {
        E item = s.get(len - 1);
        len--;
        if (len == 0) return item;
        E root = s.get(0);
        /* swap(item, data[0]); sift_down_to_bottom(0) with the hole holding `item` */
        const uint32_t end = len;
        const uint32_t lim = end >= 2 ? end - 2 : 0; /* end.saturating_sub(2) */
        uint32_t pos = 0, child = 1;
        while (child <= lim) {
            E cl = s.get(child), cr = s.get(child + 1);
            if (key(cr) <= key(cl)) { /* data[child] <= data[child+1] */
                child++;
                cl = cr;
            }
            s.set(pos, cl);
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) {
            s.set(pos, s.get(child));
            pos = child;
        }
        sift_up(s, 0, pos, item);
        return root;
    }
};

/* plain array store (shared or global) */
template <typename E>
struct ArrayStore {
    E *p;
    __device__ __forceinline__ E get(uint32_t i) const { return p[i]; }
    __device__ __forceinline__ void set(uint32_t i, E v) { p[i] = v; }
    __device__ __forceinline__ E get_sm(uint32_t i) const { return p[i]; }
    __device__ __forceinline__ void set_sm(uint32_t i, E v) { p[i] = v; }
    /* entries i and i+1, i even: one aligned load of both */
    __device__ __forceinline__ void get2(uint32_t i, E &a, E &b) const {
        if constexpr (sizeof(E) == 4) {
            uint2 v = *reinterpret_cast<const uint2 *>(p + i);
            a = v.x;
            b = v.y;
        } else {
            ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p + i);
            a = v.x;
            b = v.y;
        }
    }
    /* entries i..i+3, i a multiple of 4 */
    __device__ __forceinline__ void get4(uint32_t i, E &a, E &b, E &c, E &d) const {
        if constexpr (sizeof(E) == 4) {
            uint4 v = *reinterpret_cast<const uint4 *>(p + i);
            a = v.x;
            b = v.y;
            c = v.z;
            d = v.w;
        } else {
            ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p + i), w = *reinterpret_cast<const ulonglong2 *>(p + i + 2);
            a = v.x;
            b = v.y;
            c = w.x;
            d = w.y;
        }
    }
};

/* first `hs` entries in shared memory, the rest in the warp's HBM workspace */
template <typename E>
struct SplitStore {
    E *sm;
    E *gl;
    uint32_t hs;
    __device__ __forceinline__ E get(uint32_t i) const { return i < hs ? sm[i] : gl[i]; }
    __device__ __forceinline__ E get_sm(uint32_t i) const { return sm[i]; } /* caller knows i < hs */
    __device__ __forceinline__ void set_sm(uint32_t i, E v) { sm[i] = v; }
    __device__ __forceinline__ void get2(uint32_t i, E &a, E &b) const { /* hs is even: a pair never straddles */
        const E *q = i < hs ? sm + i : gl + i;
        a = q[0];
        b = q[1];
    }
    __device__ __forceinline__ void get4(uint32_t i, E &a, E &b, E &c, E &d) const { /* hs is a multiple of 4 */
        ArrayStore<E> q{i < hs ? sm : gl};
        q.get4(i, a, b, c, d);
    }
    __device__ __forceinline__ void set(uint32_t i, E v) {
        if (i < hs) sm[i] = v;
        else gl[i] = v;
    }
};
