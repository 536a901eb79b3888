// Minimal CUDA execution-model emulator for CPU-only kernel-logic tests.
//
// TEST INFRASTRUCTURE ONLY (think compute-sanitizer, not a backend): tests/emul/build.py compiles
// the library's .cu sources with g++ against this header into tests/emul/_build/libdplonk_emul.so
// so that indexing, shared-memory exchange, barriers and the host-side task logic can be debugged
// on a box without a GPU.  The shipped library (distributed_plonk_b200/_build/libdplonk.so) is
// nvcc-only, never includes this file, and the Python package refuses to load anything else.
//
// Model: one CUDA block at a time; blockDim threads are real std::threads; __syncthreads is a
// std::barrier; `__shared__` is plain `static` storage (one block alive at a time); dynamic shared
// memory is a per-launch heap buffer; "device" memory is host memory.
#pragma once
#include <atomic>
#include <barrier>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define DP_EMUL_ACTIVE 1

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace dp_emul {
inline thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local std::barrier<> *t_block_barrier = nullptr;
inline thread_local unsigned char *t_dyn_smem = nullptr;
struct WarpXchg {
    std::barrier<> bar;
    uint64_t slot[32];
    unsigned char blob[32][256];  // whole-struct exchange (one barrier pair per struct, not per word)
    explicit WarpXchg(int n) : bar(n) {}
};
inline thread_local WarpXchg *t_warp = nullptr;
inline thread_local unsigned t_lane = 0;

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    const unsigned nthreads = block.x * block.y * block.z;
    const unsigned nwarps = (nthreads + 31) / 32;
    std::barrier<> bar(nthreads);
    std::vector<std::unique_ptr<WarpXchg>> warps;
    for (unsigned w = 0; w < nwarps; w++) {
        unsigned n = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
        warps.emplace_back(new WarpXchg((int)n));
    }
    std::vector<unsigned char> dyn(smem + 64);
    unsigned char *dyn_aligned = (unsigned char *)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nthreads; t++) {
        pool.emplace_back([&, t] {
            t_blockDim = block;
            t_gridDim = grid;
            t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            t_block_barrier = &bar;
            t_dyn_smem = dyn_aligned;
            t_warp = warps[t / 32].get();
            t_lane = t % 32;
            for (unsigned bz = 0; bz < grid.z; bz++)
                for (unsigned by = 0; by < grid.y; by++)
                    for (unsigned bx = 0; bx < grid.x; bx++) {
                        t_blockIdx = dim3(bx, by, bz);
                        body();
                        bar.arrive_and_wait();  // next block reuses the static __shared__ storage
                    }
        });
    }
    for (auto &th : pool) th.join();
}
}  // namespace dp_emul

#define threadIdx (dp_emul::t_threadIdx)
#define blockIdx (dp_emul::t_blockIdx)
#define blockDim (dp_emul::t_blockDim)
#define gridDim (dp_emul::t_gridDim)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __constant__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

inline void __syncthreads() { dp_emul::t_block_barrier->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { dp_emul::t_warp->bar.arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <class T>
inline T dp_emul_shfl(T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    auto *w = dp_emul::t_warp;
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w->slot[dp_emul::t_lane] = raw;
    w->bar.arrive_and_wait();
    uint64_t got = w->slot[src_lane & 31];
    w->bar.arrive_and_wait();
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
// struct-granular shuffle used by kernels under DP_EMUL to keep barrier counts (thread switches) low
template <class T>
inline T dp_emul_shfl_struct(const T &v, unsigned src_lane) {
    static_assert(sizeof(T) <= 256, "shuffle blob");
    auto *w = dp_emul::t_warp;
    memcpy(w->blob[dp_emul::t_lane], &v, sizeof(T));
    w->bar.arrive_and_wait();
    T out;
    memcpy(&out, w->blob[src_lane & 31], sizeof(T));
    w->bar.arrive_and_wait();
    return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return dp_emul_shfl(v, (unsigned)src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return dp_emul_shfl(v, dp_emul::t_lane ^ (unsigned)m); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    unsigned s = dp_emul::t_lane + d;
    return dp_emul_shfl(v, s < 32 ? s : dp_emul::t_lane);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned bit = pred ? (1u << dp_emul::t_lane) : 0u, acc = 0;
    for (int l = 0; l < 32; l++) acc |= dp_emul_shfl(bit, (unsigned)l);
    return acc;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}

inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicMax(unsigned *p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
template <class T> inline T __ldg(const T *p) { return *p; }

// ------------------------------------------------------------------ runtime API subset
typedef int cudaError_t;
typedef struct dp_emul_stream *cudaStream_t;
typedef struct dp_emul_event { std::chrono::steady_clock::time_point t; } *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; size_t totalGlobalMem; char name[256]; int major, minor; };

inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 8; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof *p);
    p->multiProcessorCount = 4;
    p->totalGlobalMem = (size_t)8 << 30;
    snprintf(p->name, sizeof p->name, "cpu-emulator");
    p->major = 10;
    return cudaSuccess;
}
inline cudaError_t cudaMalloc(void **p, size_t n) {
    *p = aligned_alloc(256, (n + 255) & ~(size_t)255);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
template <class T> inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h,
                                     cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t i = 0; i < h; i++) memmove((char *)d + i * dp, (const char *)s + i * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new dp_emul_event; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new dp_emul_event; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated error"; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
// CUDA IPC: all emulated "devices" live in this process, so a handle is just the pointer
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
    memset(h, 0, sizeof *h);
    memcpy(h->reserved, &p, sizeof p);
    return cudaSuccess;
}
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
    memcpy(p, h.reserved, sizeof *p);
    return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }
