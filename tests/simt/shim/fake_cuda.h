// fake_cuda.h — the slice of the CUDA runtime API that pgvectorscale_b200/csrc/diskann_b200.cu uses, as host functions:
// "device memory" is host memory, streams and events do nothing, the one "device" reports the shape of a B200
// (SIMT_SM_COUNT overrides the SM count so that small test batches get more than one block).  With
// tests/simt/cu2cpp.py turning `kernel<<<g, b, s, st>>>(args)` into simt::launch(...), the product's whole C ABI
// builds into tests/simt/_build/libdiskann_b200_emu.so and runs on the CPU.  TEST INFRASTRUCTURE ONLY: the package
// never looks for that file; tests point the ctypes mirror at it explicitly.
#pragma once
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

struct cudaDeviceProp {
    int multiProcessorCount;
    size_t sharedMemPerBlockOptin;
};

inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { /* SIMT_FAKE_DEVICES=2: a two-GPU box for the dann_group tests */
    const char *s = getenv("SIMT_FAKE_DEVICES");
    *n = s && *s ? atoi(s) : 1;
    return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    const char *s = getenv("SIMT_SM_COUNT");
    p->multiProcessorCount = s && *s ? atoi(s) : 148;
    p->sharedMemPerBlockOptin = 232448;
    return cudaSuccess;
}
/* failure injection for the error paths: the k-th cudaMalloc from now on (0-based) reports out of memory */
inline long g_fake_malloc_countdown = -1;
/* (set through fake_cuda_fail_malloc_after(), which build_emu.py defines in the emulated ABI) */

inline cudaError_t cudaMalloc(void **p, size_t bytes) {
    if (g_fake_malloc_countdown >= 0 && g_fake_malloc_countdown-- == 0) {
        *p = nullptr;
        return cudaErrorMemoryAllocation;
    }
    const size_t sz = (bytes + 255) & ~(size_t)255;
    *p = aligned_alloc(256, sz ? sz : 256);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <typename T>
inline cudaError_t cudaMalloc(T **p, size_t bytes) {
    return cudaMalloc(reinterpret_cast<void **>(p), bytes);
}
inline cudaError_t cudaFree(void *p) {
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    if (n) memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemcpy2D(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, cudaMemcpyKind) {
    for (size_t r = 0; r < height; r++) memcpy((char *)d + r * dpitch, (const char *)s + r * spitch, width);
    return cudaSuccess;
}
/* every pointer of the fake runtime is "device memory" of device 0 */
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes {
    cudaMemoryType type;
    int device;
};
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) {
    a->type = cudaMemoryTypeDevice;
    a->device = 0;
    return cudaSuccess;
}
/* "free HBM": plenty unless DANN_FAKE_FREE_MB says otherwise (lets a test see the plan shrink its slot count) */
inline cudaError_t cudaMemGetInfo(size_t *fr, size_t *tot) {
    const char *e = getenv("DANN_FAKE_FREE_MB");
    *tot = (size_t)180 << 30;
    *fr = e && *e ? (size_t)strtoull(e, nullptr, 10) << 20 : (size_t)160 << 30;
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void *p, int v, size_t n) {
    if (n) memset(p, v, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { return cudaMemset(p, v, n); }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    *s = (void *)1;
    return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
/* failure injection: the k-th cudaStreamSynchronize from now on reports a (sticky-style) launch failure */
inline long g_fake_sync_countdown = -1;
inline cudaError_t cudaStreamSynchronize(cudaStream_t) {
    if (g_fake_sync_countdown >= 0 && g_fake_sync_countdown-- == 0) return cudaErrorInvalidValue;
    return cudaSuccess;
}
inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = (void *)1;
    return cudaSuccess;
}
enum { cudaEventDisableTiming = 2 };
#define cudaStreamLegacy ((cudaStream_t)0x1)
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) {
    *e = (void *)1;
    return cudaSuccess;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; } /* everything is in order here */
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) {
    *ms = 0.0f;
    return cudaSuccess;
}
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute attr, int value) {
    /* like the real runtime: more dynamic shared memory than the device's opt-in limit is an error, not a silent overflow */
    if (attr == cudaFuncAttributeMaxDynamicSharedMemorySize && value > 232448) return cudaErrorInvalidValue;
    return cudaSuccess;
}
template <typename F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) {
    *n = 1;
    return cudaSuccess;
}
