// topk.cuh -- the k largest estimates of every source, selected on the device.
//
// What a PPR query reads back: not the V-vector (9 MB per source on the youtube-shaped config, 333 MB on the
// Twitter-shaped one) but its head.  The reference has no counterpart -- its ValidateResult copies the whole vector
// to the host (gpu/PPRRevPushGPU.cuh:134-139); this is the read-back bench.py's end-to-end number includes.
//
// Fast path (three small kernels, p read ONCE and coalesced -- the state is vertex-major, a row holds all sources):
//   topk_threshold  per source, the k-th largest value among k or more DISTINCT vertices -- the best of the previous
//                   call (estimates move little from batch to batch), or the first vertices in internal order on the
//                   first call.  The k-th largest of any k actual values is a lower bound of the k-th largest of all.
//   topk_filter     one pass over the rows: every (vertex, source) at or above its source's bound goes on that source's
//                   candidate list (a handful more than k once the bound comes from the previous winners).
//   topk_merge      per source: sorts the candidates (bitonic, shared memory), writes the best k, remembers them.
// A source whose candidate list overflows (a loose first-call bound) gets a tighter bound -- the k-th best of the candidates
// it did collect, again k actual values -- and is filtered again; if three rounds do not settle it (a flat vector: the
// bound cannot rise past a tie) it is redone by the exact scan below.
//
// Exact scan (any distribution; strided over a vertex-major state, so it is the fallback, not the default): a CTA scans a
// slice of one source's vector, keeps candidates above a running threshold (the k-th largest value it has seen so far) in
// a shared-memory buffer, and whenever the buffer fills up sorts it, keeps the best k and raises the threshold.
// topk_merge then merges the per-slice winners.
// Order: value descending, ties by ascending CALLER vertex id (deterministic, identical on both paths).
#pragma once
#include "common.cuh"

namespace dppr {

constexpr int kTopKMax = 128;
constexpr int kTopBuf = 2048;                 // candidates a CTA sorts at a time (power of two)
constexpr int kTopItems = 4;                  // vector entries a thread inspects per round
constexpr int kTopSlice = 1 << 16;            // vector entries per CTA of the exact scan
constexpr int kTopSample = 1024;              // first-call sample: the first vertices in internal order (the highest out-degrees)
constexpr int kTopCand = 16384;               // candidate list capacity per source

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

// sorts the buffer's first next_pow2(cnt) entries (those at and beyond cnt are padded with the minimum), best first.  CTA-wide.
__device__ __forceinline__ void top_sort(TopSmem &sm) {
    __syncthreads();
    const unsigned n = sm.cnt;
    unsigned m = 2;
    while (m < n) m <<= 1;  // (<= kTopBuf: the callers compress before the buffer can overflow)
    for (unsigned i = threadIdx.x; i < m; i += kThreads)
        if (i >= n) { sm.key[i] = 0ull; sm.id[i] = 0xffffffffu; }
    __syncthreads();
    for (unsigned size = 2; size <= m; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0; stride >>= 1) {
            for (unsigned t = threadIdx.x; t < m / 2; t += kThreads) {
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

// "another round of offers may not fit": the same answer in every thread.  Call right after the barrier that follows a round of
// offers; the second barrier keeps the next round's offers (which bump sm.cnt) behind the slowest thread's read -- without
// it a thread could see the count already raised, decide alone that the buffer is full and enter top_compress's barriers
// without the others (found by racecheck in round 2; the scan had this since round 1).
__device__ __forceinline__ bool top_nearly_full(TopSmem &sm) {
    const unsigned c = sm.cnt;
    __syncthreads();
    return c > (unsigned)(kTopBuf - kThreads * kTopItems);
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
        if (top_nearly_full(sm)) tau = top_compress(sm, k);
    }
    top_compress(sm, k);
    const size_t out = ((size_t)s * gridDim.x + blockIdx.x) * (size_t)k;
    for (int i = threadIdx.x; i < k; i += kThreads) {
        const bool have = (unsigned)i < sm.cnt;
        pkey[out + i] = have ? sm.key[i] : 0ull;
        pid[out + i] = have ? sm.id[i] : 0xffffffffu;
    }
}

// one CTA per source: the best k of the source's list -- `uniform_n` entries, or counts[s] of them (clipped to `stride`; a
// longer list sets overflow[s] and leaves the result to the caller's fallback).  prev / prevk (nullable): the winners'
// INTERNAL ids, kept for the next call's bound.
__global__ void __launch_bounds__(kThreads)
    topk_merge(const unsigned long long *__restrict__ pkey, const uint32_t *__restrict__ pid, int64_t stride, int64_t uniform_n,
               unsigned int *__restrict__ counts, int k, int32_t *__restrict__ ids, double *__restrict__ values,
               const uint32_t *__restrict__ perm, uint32_t *__restrict__ prev, int *__restrict__ prevk, int *__restrict__ overflow,
               int only_overflowed, unsigned long long *__restrict__ bound) {
    __shared__ TopSmem sm;
    const int s = blockIdx.x;
    if (only_overflowed && !overflow[s]) return;  // (settled in an earlier round; uniform for the CTA)
    if (threadIdx.x == 0) sm.cnt = 0;
    __syncthreads();
    unsigned long long tau = 0ull;
    int64_t n = uniform_n;
    if (counts) {
        const unsigned int c = counts[s];
        n = c < (unsigned int)stride ? (int64_t)c : stride;
        if (threadIdx.x == 0 && overflow) overflow[s] = c > (unsigned int)stride ? (int)c : 0;  // (the count, for diagnostics)
        __syncthreads();
        if (threadIdx.x == 0) counts[s] = 0u;  // (ready for the next call)
    }
    const size_t in = (size_t)s * (size_t)stride;
    const int keep = prev ? kTopKMax : k;  // (the next call's bound wants more than k remembered: see topk_threshold)
    for (int64_t base = 0; base < n; base += kThreads * kTopItems) {
#pragma unroll
        for (int j = 0; j < kTopItems; ++j) {
            const int64_t i = base + j * kThreads + threadIdx.x;
            const bool valid = i < n && pid[in + (i < n ? i : 0)] != 0xffffffffu;
            top_offer(sm, valid, valid ? pkey[in + i] : 0ull, valid ? pid[in + i] : 0u, tau);
        }
        __syncthreads();
        if (top_nearly_full(sm)) tau = top_compress(sm, keep);
    }
    top_compress(sm, keep);
    const unsigned held = sm.cnt;
    if (threadIdx.x == 0 && bound && held >= (unsigned)k && sm.key[k - 1] > bound[s]) bound[s] = sm.key[k - 1];  // (k actual values: a valid bound for another round)
    for (int i = threadIdx.x; i < keep; i += kThreads) {
        const bool have = (unsigned)i < held;
        if (i < k) {
            ids[(size_t)s * k + i] = have ? (int32_t)sm.id[i] : -1;
            values[(size_t)s * k + i] = have ? top_unkey(sm.key[i]) : 0.0;
        }
        if (prev) prev[(size_t)s * kTopKMax + i] = have ? (perm ? perm[sm.id[i]] : sm.id[i]) : 0xffffffffu;
    }
    if (threadIdx.x == 0 && prevk) prevk[s] = (int)held;
}

// one CTA per source: bound[s] = key of HALF the k-th largest estimate among the remembered vertices (if at least k are
// held), else of the k-th largest among the first kTopSample vertices in internal order; 0 (no bound) if neither set has k
// members.  Half: a remembered vertex can lose its whole estimate in one batch (a leaf whose only edge to the source
// expires), and the candidates the looser bound admits are what the NEXT call remembers beyond its k winners, so a few
// such collapses still leave k good values (measured on BASELINE configs[3]: without the slack a fifth of the sources
// met a collapsed winner in every batch and went through two more filter rounds).
__global__ void __launch_bounds__(kThreads)
    topk_threshold(const double *__restrict__ p, int64_t Sr, int32_t V, int first_source, int k, const uint32_t *__restrict__ prev,
                   const int *__restrict__ prevk, unsigned long long *__restrict__ bound) {
    __shared__ TopSmem sm;
    const int s = blockIdx.x;
    const double *ps = p + (first_source + s);
    if (threadIdx.x == 0) sm.cnt = 0;
    __syncthreads();
    const bool use_prev = prevk[first_source + s] >= k;
    const int n = use_prev ? prevk[first_source + s] : min(kTopSample, V);
    for (int i = threadIdx.x; i < n; i += kThreads) {
        const uint32_t v = use_prev ? prev[(size_t)(first_source + s) * kTopKMax + i] : (uint32_t)i;
        const bool valid = v < (uint32_t)V;
        top_offer(sm, valid, valid ? top_key(ps[(int64_t)v * Sr]) : 0ull, v, 0ull);
    }
    __syncthreads();
    unsigned long long tau = top_compress(sm, k);
    if (use_prev && tau) {
        const double kth = top_unkey(tau);
        if (kth > 0.0) tau = top_key(0.5 * kth);
    }
    if (threadIdx.x == 0) bound[s] = tau;
}

// the rows of p, once: a warp takes 32 consecutive values of a row (or 32 / Sr whole rows when a row is shorter)
__global__ void __launch_bounds__(kThreads)
    topk_filter(const double *__restrict__ p, int64_t Sr, int32_t V, int first_source, int n_sources, const uint32_t *__restrict__ inv,
                const unsigned long long *__restrict__ bound, unsigned long long *__restrict__ ckey, uint32_t *__restrict__ cid,
                unsigned int *__restrict__ counts, const int *__restrict__ only) {
    const uint32_t lane = threadIdx.x & 31u;
    const int64_t warp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * kThreads) >> 5;
    const uint32_t rpw = Sr < 32 ? 32u / (uint32_t)Sr : 1u;  // rows per warp
    const uint32_t row_in_warp = Sr < 32 ? lane / (uint32_t)Sr : 0u, col0 = Sr < 32 ? lane % (uint32_t)Sr : lane;
    for (int64_t v0 = warp * rpw; v0 < V; v0 += nwarps * rpw) {
        const int64_t v = v0 + row_in_warp;
        if (v >= V || row_in_warp >= rpw) continue;
        for (uint32_t c = col0; c < (uint32_t)Sr; c += 32u) {
            const int s = (int)c - first_source;
            if (s < 0 || s >= n_sources || (only && !only[s])) continue;
            const unsigned long long key = top_key(__ldcs(&p[v * Sr + c]));
            if (key >= bound[s]) {
                const unsigned int pos = atomicAdd(&counts[s], 1u);
                if (pos < (unsigned int)kTopCand) {
                    ckey[(size_t)s * kTopCand + pos] = key;
                    cid[(size_t)s * kTopCand + pos] = inv ? inv[v] : (uint32_t)v;
                }
            }
        }
    }
}

}  // namespace dppr
