// repair.cuh -- residual repair for a batch of inserted / deleted edges (north-star subsystem 2).
//
// Replaces CopyOutDegree / RevertOutDegree / RevStreamUpdateOriginal (gpu/StreamUpdate.cuh:8-76),
// which applies the entries one at a time under a per-vertex atomicCAS spin lock inside a
// warp-vote retry loop -- serialising every hub and unsafe under independent thread scheduling.
//
// Each sequential step restores the push invariant
//     p[u] + a*r[u] = a*[u==s] + (1-a)/(outdeg(u)+1) * sum_{w in out(u)} p[w]
// for the intermediate edge multiset, and p is constant during the repair, so the net effect on a
// vertex u has a closed form:
//     D  = sum_{inserted (u,v)} p[v] - sum_{deleted (u,v)} p[v],  d0 / d1 = out-degree before / after
//     r'[u] = ( ((p[u] + a r[u] - a e)(d0+1) + (1-a) D) / (d1+1) - (p[u] - a e) ) / a,   e = [u==s]
// Lock-free: entries arrive sorted by u (window.cuh), the gather of p[v] is one coalesced pass with
// a warp-segmented reduction, only run partials that straddle a warp use an atomic, and one
// thread per (source, u) finalises.  tests/test_gpu_parity.py::test_repair_closed_form_matches_sequential_on_adversarial_batches
// checks this against the sequential form.
#pragma once
#include "common.cuh"
#include "window.cuh"

namespace dppr {

// S == 1: grid.x tiles the entries; one gather of p[v] per entry, runs reduced inside the warp
__global__ void __launch_bounds__(kThreads)
    repair_accumulate(const uint32_t *__restrict__ val, int64_t n, const uint32_t *__restrict__ segof,
                      const double *__restrict__ p, double *__restrict__ delta) {
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    const int64_t rounds = (n + stride - 1) / stride;
    for (int64_t rd = 0; rd < rounds; ++rd) {  // uniform trip count: every lane takes part in the shuffles
        const int64_t i = rd * stride + (int64_t)blockIdx.x * kThreads + threadIdx.x;
        const bool valid = i < n;
        uint32_t seg = 0xffffffffu;
        double x = 0.0;
        if (valid) {
            const uint32_t e = val[i];
            seg = segof[i];
            const double pv = p[e >> 1];
            x = (e & 1u) ? pv : -pv;
        }
        // inclusive segmented sum over the warp (runs are contiguous because entries are sorted)
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double y = __shfl_up_sync(kFull, x, off);
            const uint32_t sp = __shfl_up_sync(kFull, seg, off);
            if ((int)lane_id() >= off && sp == seg) x += y;
        }
        const uint32_t snext = __shfl_down_sync(kFull, seg, 1);
        const bool tail = (lane_id() == 31) || (snext != seg);
        if (valid && tail) atomicAdd(&delta[seg], x);
    }
}

// S > 1 (state is vertex-major: the S values of a vertex are one contiguous row): a warp takes 32 consecutive entries and
// walks them in order; its lanes are SOURCES (grid.y = block of 128 sources, 4 per lane), so every gather is a coalesced
// read of a row of p, and a run's partial sum leaves the registers once per run per warp -- one coalesced row of atomics.
constexpr int kRepairCols = 4;  // sources per lane
__global__ void __launch_bounds__(kThreads)
    repair_accumulate_rows(const uint32_t *__restrict__ val, int64_t n, const uint32_t *__restrict__ segof,
                           const double *__restrict__ p, int64_t Sr, int S, double *__restrict__ delta) {
    const int s_base = blockIdx.y * 32 * kRepairCols;
    const int64_t warps = (int64_t)gridDim.x * kWarps;
    for (int64_t base = ((int64_t)blockIdx.x * kWarps + warp_id()) * 32; base < n; base += warps * 32) {
        const int64_t i = base + lane_id();
        const uint32_t my_e = i < n ? val[i] : 0u, my_seg = i < n ? segof[i] : 0xffffffffu;
        const int cnt = (int)min((int64_t)32, n - base);
        double acc[kRepairCols];
#pragma unroll
        for (int j = 0; j < kRepairCols; ++j) acc[j] = 0.0;
        uint32_t cur = __shfl_sync(kFull, my_seg, 0);
        for (int t = 0; t < cnt; ++t) {
            const uint32_t e = __shfl_sync(kFull, my_e, t), seg = __shfl_sync(kFull, my_seg, t);
            if (seg != cur) {  // (warp-uniform)
#pragma unroll
                for (int j = 0; j < kRepairCols; ++j) {
                    const int s = s_base + j * 32 + (int)lane_id();
                    if (s < S && acc[j] != 0.0) atomicAdd(&delta[(int64_t)cur * Sr + s], acc[j]);
                    acc[j] = 0.0;
                }
                cur = seg;
            }
            const double *row = p + (int64_t)(e >> 1) * Sr;
#pragma unroll
            for (int j = 0; j < kRepairCols; ++j) {
                const int s = s_base + j * 32 + (int)lane_id();
                if (s < S) {
                    const double pv = row[s];
                    acc[j] += (e & 1u) ? pv : -pv;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kRepairCols; ++j) {
            const int s = s_base + j * 32 + (int)lane_id();
            if (s < S && acc[j] != 0.0) atomicAdd(&delta[(int64_t)cur * Sr + s], acc[j]);
        }
    }
}

// one thread per (run, source), sources fastest; also clears delta for the next batch
__global__ void __launch_bounds__(kThreads)
    repair_finalize(Segments sg, const int32_t *__restrict__ seg_d0, const int32_t *__restrict__ src, int S, int64_t Sr,
                    const double *__restrict__ p, double *__restrict__ r, double *__restrict__ delta, double alpha) {
    const uint32_t nseg = *sg.count;
    const int64_t total = (int64_t)nseg * S;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
        const uint32_t g = (uint32_t)(t / S);
        const int s = (int)(t - (int64_t)g * S);
        const uint32_t u = sg.vertex[g];
        const uint32_t st = sg.start[g], fi = sg.first_ins[g], en = sg.start[g + 1];
        const double d0 = (double)seg_d0[g];
        const double d1 = d0 + (double)(en - fi) - (double)(fi - st);
        const int64_t idx = (int64_t)u * Sr + s;
        const double D = delta[(int64_t)g * Sr + s];
        delta[(int64_t)g * Sr + s] = 0.0;
        const double ae = (src[s] == (int32_t)u) ? alpha : 0.0;
        const double pu = p[idx], ru = r[idx];
        r[idx] = (((pu + alpha * ru - ae) * (d0 + 1.0) + (1.0 - alpha) * D) / (d1 + 1.0) - (pu - ae)) / alpha;
    }
}

}  // namespace dppr
