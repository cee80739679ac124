// common.cuh -- shared helpers for the sm_100a kernels (no CUB / Thrust / cuSPARSE anywhere).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace dppr {

constexpr int kThreads = 256;          // CTA size used by every kernel
constexpr int kWarps = kThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

struct CudaFailure : std::runtime_error {
    explicit CudaFailure(const std::string &m) : std::runtime_error(m) {}
};
struct InvalidArgument : std::runtime_error {
    explicit InvalidArgument(const std::string &m) : std::runtime_error(m) {}
};
struct StateError : std::runtime_error {
    explicit StateError(const std::string &m) : std::runtime_error(m) {}
};
struct CapacityError : std::runtime_error {
    explicit CapacityError(const std::string &m) : std::runtime_error(m) {}
};

// The reference prints and exit(-1)s on a CUDA error (gpu/GPUUtil.cuh:7-19); a library must not,
// so failures become exceptions that capi.cu turns into DPPR_E_CUDA + dppr_last_error().
#define DPPR_CUDA(call)                                                                               \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess)                                                                     \
            throw ::dppr::CudaFailure(std::string(#call) + " failed at " + __FILE__ + ":" +           \
                                      std::to_string(__LINE__) + ": " + cudaGetErrorString(err__));   \
    } while (0)

// number of kernels this library has launched from the calling thread (bench.py reports it)
inline unsigned long long &launch_counter() {
    static thread_local unsigned long long c = 0;
    return c;
}

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// owning device buffer
template <typename T>
struct DevBuf {
    T *ptr = nullptr;
    size_t count = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void alloc(size_t n) {
        release();
        count = n;
        if (n) DPPR_CUDA(cudaMalloc((void **)&ptr, n * sizeof(T)));
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        count = 0;
    }
    size_t bytes() const { return count * sizeof(T); }
};

template <typename T>
struct PinnedBuf {
    T *ptr = nullptr;
    size_t count = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    ~PinnedBuf() { release(); }
    void alloc(size_t n) {
        release();
        count = n;
        if (n) DPPR_CUDA(cudaMallocHost((void **)&ptr, n * sizeof(T)));
    }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        count = 0;
    }
};

// ---- device-side building blocks ------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// inclusive warp scan by shuffles
template <typename T>
__device__ __forceinline__ T warp_inclusive_sum(T x) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        T y = __shfl_up_sync(kFull, x, off);
        if ((int)lane_id() >= off) x += y;
    }
    return x;
}

// CTA-wide exclusive scan of one value per thread (kThreads threads); every thread gets its
// exclusive prefix, `total` the CTA sum.  `smem` needs kWarps + 1 elements.  Three barriers (the last one frees smem).
template <typename T>
__device__ __forceinline__ T block_exclusive_sum(T x, T *smem, T &total) {
    T inc = warp_inclusive_sum(x);
    if (lane_id() == 31) smem[warp_id()] = inc;
    __syncthreads();
    if (warp_id() == 0) {
        T w = (lane_id() < kWarps) ? smem[lane_id()] : T(0);
        T winc = warp_inclusive_sum(w);
        if (lane_id() < kWarps) smem[lane_id()] = winc - w;  // exclusive prefix of warp sums
        if (lane_id() == kWarps - 1) smem[kWarps] = winc;
    }
    __syncthreads();
    T res = smem[warp_id()] + inc - x;
    total = smem[kWarps];
    __syncthreads();  // smem reusable on return
    return res;
}

}  // namespace dppr
