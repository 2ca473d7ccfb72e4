// Stand-in for cub::DeviceRadixSort::SortPairs under the CPU SIMT emulator: a stable sort on the key bits
// [begin_bit, end_bit) (a radix sort is stable), with cub's two-call temp-storage protocol.  Test infrastructure.
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

namespace cub {
struct DeviceRadixSort {
    template <typename K, typename V>
    static cudaError_t SortPairs(void *d_temp, size_t &temp_bytes, const K *keys_in, K *keys_out, const V *vals_in, V *vals_out,
                                 size_t n, int begin_bit, int end_bit, cudaStream_t = nullptr) {
        if (!d_temp) {
            temp_bytes = 256;
            return cudaSuccess;
        }
        std::vector<size_t> idx(n);
        std::iota(idx.begin(), idx.end(), (size_t)0);
        const int width = end_bit - begin_bit;
        const K mask = width >= (int)(8 * sizeof(K)) ? ~K(0) : ((K(1) << width) - K(1));
        auto key = [&](size_t i) { return (K)((keys_in[i] >> begin_bit) & mask); };
        std::stable_sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return key(a) < key(b); });
        for (size_t i = 0; i < n; i++) {
            keys_out[i] = keys_in[idx[i]];
            vals_out[i] = vals_in[idx[i]];
        }
        return cudaSuccess;
    }
};
}  // namespace cub
