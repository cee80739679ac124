// window_fused.cuh -- the whole per-batch window update in ONE single-CTA launch, for small batches.
//
// The multi-kernel path of window.cuh / primitives.cuh costs ~24 launches per batch (entries, 3 x (histogram,
// scan, scatter), run-length encoding, plan, relocate, insert, out-degrees); at BASELINE configs[1]
// (11,948 entries per batch) each kernel runs for ~2 us and the sequence is purely launch-bound: 0.105 ms per
// batch (profiles/launches_r01a_youtube_3batches.csv).  For SMALL batches (mode 1, -c 100: BASELINE configs[2])
// one CTA of 1024 threads does the same work with __syncthreads() where the kernel boundaries were; the
// per-element bodies are the same device functions the multi-kernel path uses, so both produce identical graphs.
#pragma once
#include "window.cuh"
#include "repair.cuh"

namespace dppr {

constexpr int kFusedThreads = 1024;
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kFusedMaxEntries = 2048;    // measured: 200 entries 0.025 ms (multi-kernel 0.116); 11,948 entries 0.160 ms
                                          // (multi-kernel 0.099): one CTA lacks the memory-level parallelism beyond ~2k

struct FusedSmem {
    uint32_t wcnt[kFusedWarps][256];  // per-warp digit counts / running offsets
    uint32_t dbase[256];              // exclusive prefix of the digit totals
    uint32_t wsum[kFusedWarps + 1];
};

// exclusive prefix over the 1024 threads of the CTA
__device__ __forceinline__ uint32_t fused_exclusive_sum(uint32_t x, uint32_t *wsum, uint32_t &total) {
    const uint32_t inc = warp_inclusive_sum(x);
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        const uint32_t v = wsum[l];
        const uint32_t vi = warp_inclusive_sum(v);
        wsum[l] = vi - v;
        if (l == 31) wsum[kFusedWarps] = vi;
    }
    __syncthreads();
    const uint32_t res = wsum[w] + inc - x;
    total = wsum[kFusedWarps];
    __syncthreads();
    return res;
}

// stable LSD radix sort of n pairs inside one CTA; returns 0 / 1 = result in (k0, v0) / (k1, v1)
__device__ int fused_sort_pairs(uint32_t *k0, uint32_t *v0, uint32_t *k1, uint32_t *v1, uint32_t n, int key_bits,
                                FusedSmem &sm) {
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    const uint32_t run = ((n + kFusedWarps - 1) / kFusedWarps + 31u) & ~31u;  // keys per warp, multiple of 32
    const uint32_t lo = w * run, hi = min(n, lo + run);
    const int passes = (key_bits + 7) / 8;
    uint32_t *kin = k0, *vin = v0, *kout = k1, *vout = v1;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        for (int i = threadIdx.x; i < kFusedWarps * 256; i += kFusedThreads) (&sm.wcnt[0][0])[i] = 0;
        __syncthreads();
        // walk 1: per-warp digit counts
        for (uint32_t c = lo; c < hi; c += 32) {
            const uint32_t i = c + l;
            const bool valid = i < hi;
            const uint32_t d = valid ? ((kin[i] >> shift) & 255u) : 256u;
            const unsigned peers = __match_any_sync(kFull, d);
            if (valid && l == (unsigned)(__ffs(peers) - 1)) sm.wcnt[w][d] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        // digit totals -> exclusive prefix; per (digit, warp) running offsets
        uint32_t tot = 0;
        if (threadIdx.x < 256) {
            for (int ww = 0; ww < kFusedWarps; ++ww) {
                const uint32_t c = sm.wcnt[ww][threadIdx.x];
                sm.wcnt[ww][threadIdx.x] = tot;
                tot += c;
            }
        }
        uint32_t total;
        const uint32_t ex = fused_exclusive_sum(threadIdx.x < 256 ? tot : 0u, sm.wsum, total);
        if (threadIdx.x < 256) sm.dbase[threadIdx.x] = ex;
        __syncthreads();
        // walk 2: stable scatter
        for (uint32_t c = lo; c < hi; c += 32) {
            const uint32_t i = c + l;
            const bool valid = i < hi;
            uint32_t key = 0, val = 0, d = 256u;
            if (valid) { key = kin[i]; val = vin[i]; d = (key >> shift) & 255u; }
            const unsigned peers = __match_any_sync(kFull, d);
            const uint32_t before = valid ? sm.wcnt[w][d] : 0u;
            __syncwarp();
            if (valid) {
                const uint32_t pos = sm.dbase[d] + before + __popc(peers & lanemask_lt());
                kout[pos] = key;
                vout[pos] = val;
                if (l == (unsigned)(__ffs(peers) - 1)) sm.wcnt[w][d] = before + __popc(peers);
            }
            __syncwarp();
        }
        __syncthreads();
        uint32_t *t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    return passes & 1;
}

// run-length encoding of sorted entries (same outputs as rle_heads + scan + rle_fill)
__device__ void fused_rle(const uint32_t *key, const uint32_t *val, uint32_t n, const Segments &sg, FusedSmem &sm) {
    const uint32_t per = (n + kFusedThreads - 1) / kFusedThreads;
    const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
    uint32_t heads = 0;
    for (uint32_t i = lo; i < hi; ++i) heads += (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
    uint32_t total;
    uint32_t s = fused_exclusive_sum(heads, sm.wsum, total);  // runs that start before this thread's block
    for (uint32_t i = lo; i < hi; ++i) {
        const bool head = (i == 0) || (key[i] != key[i - 1]);
        if (head) ++s;
        rle_fill_one(i, s - 1u, head, key, val, n, sg);
    }
    if (n == 0 && threadIdx.x == 0) *sg.count = 0;
    __syncthreads();
}

struct FusedArgs {
    int2 *log;
    int64_t W, log_start;
    const int2 *arriving;
    int64_t B;
    int directed;
    int key_bits;
    uint32_t *akey[2], *aval[2], *bkey[2], *bval[2];
    Segments segA, segB;
    WindowView w;
    uint32_t *ins_pos;
    RelocJob *jobs;
    uint32_t *njobs;
    int32_t *seg_d0;
    const uint32_t *perm;    // caller id -> internal id (null = identity)
    // out-lists of a directed graph (null vmeta = not maintained)
    WindowView wo;
    uint32_t *ins_posB;
    RelocJob *jobsB;
    uint32_t *njobsB;
    const uint2 *pend_prev;        // ranges released by the previous batch (window.cuh, PoolFree)
    const uint32_t *npend_prev;
};

__global__ void __launch_bounds__(kFusedThreads, 1) win_fused_small(const FusedArgs a) {
    __shared__ FusedSmem sm;
    const uint32_t B = (uint32_t)a.B;
    const uint32_t n = a.directed ? 2u * B : 4u * B;
    for (uint32_t i = threadIdx.x; i < B; i += kFusedThreads)
        batch_entries_one(i, a.log, a.W, a.log_start, a.arriving, a.B, a.directed, a.w.V, a.akey[0], a.aval[0], a.bkey[0],
                          a.bval[0], a.w.errflags, a.perm);
    {
        const uint32_t np = *a.npend_prev;
        for (uint32_t j = threadIdx.x; j < np; j += kFusedThreads) pool_reclaim_one(j, a.pend_prev, a.w.fr);
    }
    __syncthreads();
    // group A: keyed by destination -> in-lists
    const int ra = fused_sort_pairs(a.akey[0], a.aval[0], a.akey[1], a.aval[1], n, a.key_bits, sm);
    const uint32_t *ka = a.akey[ra], *va = a.aval[ra];
    fused_rle(ka, va, n, a.segA, sm);
    const uint32_t nsegA = *a.segA.count;
    for (uint32_t s = threadIdx.x; s < nsegA; s += kFusedThreads) plan_one(s, a.segA, a.w, a.ins_pos, a.jobs, a.njobs);
    __syncthreads();
    const uint32_t nj = *a.njobs;
    for (uint32_t j = threadIdx.x >> 5; j < nj; j += kFusedWarps) {  // one warp per relocated ring
        const RelocJob jb = a.jobs[j];
        for (uint32_t k = threadIdx.x & 31; k < jb.len; k += 32)
            a.w.pool[jb.new_base + k] = a.w.pool[jb.old_base + ((jb.old_head + k) & (jb.old_cap - 1u))];
    }
    for (uint32_t i = threadIdx.x; i < n; i += kFusedThreads) insert_one(i, ka, va, a.segA, a.ins_pos, a.w);
    __syncthreads();
    // group B: keyed by source -> out-degrees (+ residual repair, launched separately)
    int rb = ra;
    if (a.directed) {
        rb = fused_sort_pairs(a.bkey[0], a.bval[0], a.bkey[1], a.bval[1], n, a.key_bits, sm);
        fused_rle(a.bkey[rb], a.bval[rb], n, a.segB, sm);
    }
    const uint32_t nsegB = *a.segB.count;
    for (uint32_t s = threadIdx.x; s < nsegB; s += kFusedThreads) out_degree_one(s, a.segB, a.w.outdeg, a.seg_d0);
    if (a.directed && a.wo.vmeta != nullptr) {  // out-lists: same steps on the source-sorted entries
        if (threadIdx.x == 0) *a.njobsB = 0;
        __syncthreads();
        for (uint32_t s = threadIdx.x; s < nsegB; s += kFusedThreads) plan_one(s, a.segB, a.wo, a.ins_posB, a.jobsB, a.njobsB);
        __syncthreads();
        const uint32_t njb = *a.njobsB;
        for (uint32_t j = threadIdx.x >> 5; j < njb; j += kFusedWarps) {
            const RelocJob jb = a.jobsB[j];
            for (uint32_t k = threadIdx.x & 31; k < jb.len; k += 32)
                a.w.pool[jb.new_base + k] = a.w.pool[jb.old_base + ((jb.old_head + k) & (jb.old_cap - 1u))];
        }
        for (uint32_t i = threadIdx.x; i < n; i += kFusedThreads) insert_one(i, a.bkey[rb], a.bval[rb], a.segB, a.ins_posB, a.wo);
    }
}

}  // namespace dppr
