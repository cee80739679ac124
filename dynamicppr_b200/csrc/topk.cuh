// topk.cuh -- the k largest estimates of every source, selected on the device.
//
// What a PPR query reads back: not the V-vector (9 MB per source on the youtube-shaped config, 333 MB on the
// Twitter-shaped one) but its head.  The reference has no counterpart -- its ValidateResult copies the whole vector
// to the host (gpu/PPRRevPushGPU.cuh:134-139); this is the read-back bench.py's end-to-end number includes.
//
// One pass over p: a CTA scans a slice of one source's vector, keeps candidates above a running threshold (the k-th
// largest value it has seen so far) in a shared-memory buffer, and whenever the buffer fills up sorts it (bitonic),
// keeps the best k and raises the threshold.  A second kernel merges the per-slice winners of a source the same way.
// Order: value descending, ties by ascending CALLER vertex id (deterministic).
#pragma once
#include "common.cuh"

namespace dppr {

constexpr int kTopKMax = 128;
constexpr int kTopBuf = 2048;                 // candidates a CTA sorts at a time (power of two)
constexpr int kTopItems = 4;                  // vector entries a thread inspects per round
constexpr int kTopSlice = 1 << 16;            // vector entries per CTA of the first pass

struct TopSmem {
    unsigned long long key[kTopBuf];
    uint32_t id[kTopBuf];
    unsigned cnt;
};

// doubles -> unsigned keys with the same order
__device__ __forceinline__ unsigned long long top_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double top_unkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
// "a ranks before b": larger value first, then smaller id
__device__ __forceinline__ bool top_before(unsigned long long ka, uint32_t ia, unsigned long long kb, uint32_t ib) {
    return ka > kb || (ka == kb && ia < ib);
}

// sorts the whole buffer (entries at and beyond cnt are padded with the minimum), best first.  CTA-wide.
__device__ __forceinline__ void top_sort(TopSmem &sm) {
    __syncthreads();
    const unsigned n = sm.cnt;
    for (unsigned i = threadIdx.x; i < (unsigned)kTopBuf; i += kThreads)
        if (i >= n) { sm.key[i] = 0ull; sm.id[i] = 0xffffffffu; }
    __syncthreads();
    for (unsigned size = 2; size <= (unsigned)kTopBuf; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = threadIdx.x; t < (unsigned)kTopBuf / 2; t += kThreads) {
                const unsigned lo = 2 * t - (t & (stride - 1));  // index with the `stride` bit clear
                const unsigned hi = lo + stride;
                const bool desc = (lo & size) == 0;  // this sub-sequence puts the better element first
                const unsigned long long ka = sm.key[lo], kb = sm.key[hi];
                const uint32_t ia = sm.id[lo], ib = sm.id[hi];
                if (desc ? top_before(kb, ib, ka, ia) : top_before(ka, ia, kb, ib)) {
                    sm.key[lo] = kb; sm.id[lo] = ib;
                    sm.key[hi] = ka; sm.id[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
}

// keep the best k of the buffer; returns the new threshold (key of the k-th best, 0 if fewer than k are held)
__device__ __forceinline__ unsigned long long top_compress(TopSmem &sm, int k) {
    top_sort(sm);
    const unsigned n = sm.cnt;
    __syncthreads();
    if (threadIdx.x == 0) sm.cnt = n < (unsigned)k ? n : (unsigned)k;
    __syncthreads();
    return n >= (unsigned)k ? sm.key[k - 1] : 0ull;
}

// candidates strictly above the threshold -- or equal to it (ties are resolved by id in the sort) -- enter the buffer
__device__ __forceinline__ void top_offer(TopSmem &sm, bool valid, unsigned long long key, uint32_t id, unsigned long long tau) {
    if (valid && key >= tau) {
        const unsigned pos = atomicAdd(&sm.cnt, 1u);
        sm.key[pos] = key;
        sm.id[pos] = id;
    }
}

// grid (slices, sources).  partial: [source][slice][k] keys + ids, short lists padded with key 0.
__global__ void __launch_bounds__(kThreads)
    topk_partial(const double *__restrict__ p, int64_t Sr, int32_t V, int first_source, const uint32_t *__restrict__ inv, int k,
                 unsigned long long *__restrict__ pkey, uint32_t *__restrict__ pid) {
    __shared__ TopSmem sm;
    const int s = blockIdx.y;
    const double *ps = p + (first_source + s);  // state is vertex-major: element v of this source sits at v * Sr
    if (threadIdx.x == 0) sm.cnt = 0;
    __syncthreads();
    unsigned long long tau = 0ull;
    const int64_t lo = (int64_t)blockIdx.x * kTopSlice, hi = min((int64_t)V, lo + kTopSlice);
    for (int64_t base = lo; base < hi; base += kThreads * kTopItems) {
#pragma unroll
        for (int j = 0; j < kTopItems; ++j) {
            const int64_t v = base + j * kThreads + threadIdx.x;
            const bool valid = v < hi;
            const unsigned long long key = valid ? top_key(ps[v * Sr]) : 0ull;
            top_offer(sm, valid, key, valid ? (inv ? inv[v] : (uint32_t)v) : 0u, tau);
        }
        __syncthreads();
        if (sm.cnt > (unsigned)(kTopBuf - kThreads * kTopItems)) tau = top_compress(sm, k);  // (uniform)
    }
    top_compress(sm, k);
    const size_t out = ((size_t)s * gridDim.x + blockIdx.x) * (size_t)k;
    for (int i = threadIdx.x; i < k; i += kThreads) {
        const bool have = (unsigned)i < sm.cnt;
        pkey[out + i] = have ? sm.key[i] : 0ull;
        pid[out + i] = have ? sm.id[i] : 0xffffffffu;
    }
}

// one CTA per source: merge `slices` lists of k
__global__ void __launch_bounds__(kThreads)
    topk_merge(const unsigned long long *__restrict__ pkey, const uint32_t *__restrict__ pid, int slices, int k,
               int32_t *__restrict__ ids, double *__restrict__ values) {
    __shared__ TopSmem sm;
    const int s = blockIdx.x;
    if (threadIdx.x == 0) sm.cnt = 0;
    __syncthreads();
    unsigned long long tau = 0ull;
    const int64_t n = (int64_t)slices * k;
    const size_t in = (size_t)s * (size_t)n;
    for (int64_t base = 0; base < n; base += kThreads * kTopItems) {
#pragma unroll
        for (int j = 0; j < kTopItems; ++j) {
            const int64_t i = base + j * kThreads + threadIdx.x;
            const bool valid = i < n && pid[in + (i < n ? i : 0)] != 0xffffffffu;
            top_offer(sm, valid, valid ? pkey[in + i] : 0ull, valid ? pid[in + i] : 0u, tau);
        }
        __syncthreads();
        if (sm.cnt > (unsigned)(kTopBuf - kThreads * kTopItems)) tau = top_compress(sm, k);
    }
    top_compress(sm, k);
    for (int i = threadIdx.x; i < k; i += kThreads) {
        const bool have = (unsigned)i < sm.cnt;
        ids[(size_t)s * k + i] = have ? (int32_t)sm.id[i] : -1;
        values[(size_t)s * k + i] = have ? top_unkey(sm.key[i]) : 0.0;
    }
}

}  // namespace dppr
