// validate.cuh -- the reference's -DVALIDATE checks, run on the device.
//
// Reference: PPRRevPushGPU::ValidateResult (gpu/PPRRevPushGPU.cuh:133-165) copies pagerank / residual to the host,
// asserts |r[u]| < eps for every u and compares the estimates with a host power iteration; ValidateGraph
// (gpu/PPRRevPushGPU.cuh:45-90) copies the device in-CSR back and compares it with a graph rebuilt from the file.  At
// Twitter scale (41.7 M vertices, 146.8 M window edges) neither copy is something to do per batch, so here:
//   * residual bound: max_u |r[u]|, one reduction over the source's residual vector;
//   * instead of a power iteration, the PUSH INVARIANT every operation of the algorithm preserves,
//         p[u] + a r[u] = a [u == s] + (1 - a) / (outdeg(u) + 1) * sum_{w in out(u)} p[w]        for every u:
//     its largest defect is computed from the device-resident window graph itself.  Together with the residual bound it
//     implies |p - pi| <= max |r| (the fixed point of the invariant with r = 0 is pi), i.e. the 2 eps parity criterion;
//   * window graph: the caller hands the W window edges (device memory, stream order), the expected canonical entry list
//     is built from them by the same stable sorts and compared entry by entry with the engine's own, on the device.
#pragma once
#include "common.cuh"
#include "window.cuh"

namespace dppr {

__device__ __forceinline__ void atomic_max_double_bits(unsigned long long *dst, double x) {  // x >= 0
    atomicMax(dst, (unsigned long long)__double_as_longlong(x));
}

// (p, r point at the source's column of the vertex-major state; Sr = row stride)
__global__ void __launch_bounds__(kThreads) val_residual_max(const double *__restrict__ r, int64_t Sr, int32_t V, unsigned long long *out) {
    double mx = 0.0;
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < V; v += (int64_t)gridDim.x * kThreads) {
        const double x = fabs(r[v * Sr]);
        mx = (x > mx || x != x) ? x : mx;  // (a NaN wins: it must not pass)
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double y = __shfl_xor_sync(kFull, mx, off);
        mx = (y > mx || y != y) ? y : mx;
    }
    if (lane_id() == 0) atomic_max_double_bits(out, mx != mx ? __longlong_as_double(0x7ff0000000000000ll) : mx);
}

// acc[u] += p[w] for every window edge u -> w: the edge sits in w's in-list as `u`.  One warp per vertex w.
__global__ void __launch_bounds__(kThreads)
    val_out_sums(const uint4 *__restrict__ vmeta, const int32_t *__restrict__ pool, const double *__restrict__ p, int64_t Sr, int32_t V,
                 double *__restrict__ acc) {
    const int64_t warps = (int64_t)gridDim.x * kWarps;
    for (int64_t w = (int64_t)blockIdx.x * kWarps + warp_id(); w < V; w += warps) {
        const uint4 m = vmeta[w];
        const double pw = p[w * Sr];
        if (pw == 0.0) continue;
        for (uint32_t k = lane_id(); k < m.z; k += 32) atomicAdd(&acc[pool[m.x + ((m.y + k) & (m.w - 1u))]], pw);
    }
}

__global__ void __launch_bounds__(kThreads)
    val_invariant(const double *__restrict__ p, const double *__restrict__ r, int64_t Sr, const int32_t *__restrict__ outdeg,
                  const double *__restrict__ acc, int32_t V, int32_t source, double alpha, unsigned long long *out) {
    double mx = 0.0;
    for (int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x; u < V; u += (int64_t)gridDim.x * kThreads) {
        const double lhs = p[u * Sr] + alpha * r[u * Sr] - (u == source ? alpha : 0.0);
        const double d = fabs(lhs - (1.0 - alpha) * acc[u] / ((double)outdeg[u] + 1.0));
        mx = (d > mx || d != d) ? d : mx;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double y = __shfl_xor_sync(kFull, mx, off);
        mx = (y > mx || y != y) ? y : mx;
    }
    if (lane_id() == 0) atomic_max_double_bits(out, mx != mx ? __longlong_as_double(0x7ff0000000000000ll) : mx);
}

// expected entries of a window given as W (src, dst) pairs in caller ids: (key = dst, val = src), mirrored if undirected
__global__ void __launch_bounds__(kThreads)
    val_window_entries(const int2 *__restrict__ pairs, int64_t W, int directed, int32_t V, uint32_t *__restrict__ key,
                       uint32_t *__restrict__ val, uint32_t *__restrict__ outdeg, unsigned long long *bad) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < W; i += (int64_t)gridDim.x * kThreads) {
        int2 e = pairs[i];
        if ((uint32_t)e.x >= (uint32_t)V || (uint32_t)e.y >= (uint32_t)V) { atomicAdd(bad, 1ull); e.x = 0; e.y = 0; }
        if (directed) {
            key[i] = (uint32_t)e.y; val[i] = (uint32_t)e.x;
            atomicAdd(&outdeg[e.x], 1u);
        } else {
            key[2 * i] = (uint32_t)e.y; val[2 * i] = (uint32_t)e.x;
            key[2 * i + 1] = (uint32_t)e.x; val[2 * i + 1] = (uint32_t)e.y;
            atomicAdd(&outdeg[e.x], 1u); atomicAdd(&outdeg[e.y], 1u);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
    val_count_diff(const uint32_t *__restrict__ a0, const uint32_t *__restrict__ b0, const uint32_t *__restrict__ a1,
                   const uint32_t *__restrict__ b1, int64_t n, unsigned long long *out) {
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        bad += (a0[i] != b0[i] || (a1 && a1[i] != b1[i])) ? 1ull : 0ull;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) bad += __shfl_xor_sync(kFull, bad, off);
    if (lane_id() == 0 && bad) atomicAdd(out, bad);
}

}  // namespace dppr
