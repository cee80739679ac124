// window_coop.cuh -- the per-batch window update as ONE cooperative launch for mid-size batches.
//
// The multi-kernel path (window.cuh + primitives.cuh) is 24 launches per batch; at BASELINE configs[1]
// (11,948 entries) that is launch-bound: 0.10 ms per batch, 8 % of the step.  Here the same device functions run
// inside one cooperative kernel, with a software grid barrier where the kernel boundaries were (~1 us for <= 64
// CTAs instead of ~4 us per launch).  The barrier's trailing ld.acquire invalidates the SM's L1, so data written
// by other CTAs in an earlier stage is re-read from L2.
#pragma once
#include "window.cuh"
#include "primitives.cuh"
#include "push.cuh"   // grid barrier helpers

namespace dppr {

constexpr int kCoopMaxTiles = 64;                         // 64 x 2048 = 131,072 entries per group
constexpr int kCoopMaxEntries = kCoopMaxTiles * kSortTile;

struct CoopArgs {
    int2 *log;
    int64_t W, log_start;
    const int2 *arriving;
    int64_t B;
    int directed;
    int key_bits;
    uint32_t *akey[2], *aval[2], *bkey[2], *bval[2];
    uint32_t *hist;          // kRadix * tiles
    uint32_t *tile_heads;    // tiles
    Segments segA, segB;
    WindowView w;
    uint32_t *ins_pos;
    RelocJob *jobs;
    uint32_t *njobs;
    int32_t *seg_d0;
    const uint32_t *perm;    // caller id -> internal id (null = identity)
    unsigned *bar;           // grid barrier counter, zero at launch
    // out-lists of a directed graph (null vmeta = not maintained)
    WindowView wo;
    uint32_t *ins_posB;
    RelocJob *jobsB;
    uint32_t *njobsB;
    const uint2 *pend_prev;        // ranges released by the previous batch (window.cuh, PoolFree)
    const uint32_t *npend_prev;
};

struct CoopSmem {
    RadixScatterSmem scatter;
    uint32_t hist[kRadix];
    uint32_t scan[kWarps + 1];
    uint32_t tile_base;
    int abort_flag;
};

__device__ __forceinline__ bool coop_barrier(unsigned *bar, unsigned &gen, int &abort_flag, int *errflags) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++gen;
        const unsigned target = gen * gridDim.x;
        bar_arrive_release(bar);
        const long long t0 = clock64();
        bool ok = true;
        while (bar_load_relaxed(bar) < target) {
            if (clock64() - t0 > 4000000000ll) { atomicOr(errflags, kErrWatchdog); ok = false; break; }
        }
        (void)bar_load_acquire(bar);
        abort_flag = ok ? 0 : 1;
    }
    __syncthreads();
    return abort_flag == 0;
}

// in-place exclusive scan of `n` uint32 by ONE CTA (n <= kRadix * kCoopMaxTiles)
__device__ void coop_scan_one_cta(uint32_t *data, uint32_t n, uint32_t *scan_sm) {
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += kThreads * 8) {
        uint32_t v[8], acc = 0;
        const uint32_t i0 = base + threadIdx.x * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) { v[k] = (i0 + k < n) ? data[i0 + k] : 0u; acc += v[k]; }
        uint32_t total;
        uint32_t run = block_exclusive_sum<uint32_t>(acc, scan_sm, total) + carry;
#pragma unroll
        for (int k = 0; k < 8; ++k) { if (i0 + k < n) data[i0 + k] = run; run += v[k]; }
        carry += total;
    }
    __syncthreads();
}

// one group: stable sort by vertex, then run-length encoding.  Returns the buffer index of the sorted pairs.
__device__ int coop_sort_and_rle(const CoopArgs &a, uint32_t *const key[2], uint32_t *const val[2], uint32_t n,
                                 const Segments &sg, CoopSmem &sm, unsigned &gen, bool &alive) {
    const int tiles = (int)((n + kSortTile - 1) / kSortTile);
    const int passes = (a.key_bits + 7) / 8;
    int cur = 0;
    for (int p = 0; p < passes && alive; ++p) {
        const int shift = 8 * p;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) radix_hist_tile(key[cur], a.hist, n, shift, tiles, t, sm.hist);
        if (!(alive = coop_barrier(a.bar, gen, sm.abort_flag, a.w.errflags))) break;
        if (blockIdx.x == 0) coop_scan_one_cta(a.hist, (uint32_t)(kRadix * tiles), sm.scan);
        if (!(alive = coop_barrier(a.bar, gen, sm.abort_flag, a.w.errflags))) break;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x)
            radix_scatter_tile(key[cur], val[cur], key[cur ^ 1], val[cur ^ 1], a.hist, n, shift, tiles, t, sm.scatter);
        if (!(alive = coop_barrier(a.bar, gen, sm.abort_flag, a.w.errflags))) break;
        cur ^= 1;
    }
    if (!alive) return cur;
    const uint32_t *k = key[cur], *v = val[cur];
    // run heads per tile
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        uint32_t cnt = 0;
        const uint32_t i0 = (uint32_t)t * kSortTile + threadIdx.x * kSortChunks;
#pragma unroll
        for (int c = 0; c < kSortChunks; ++c) {
            const uint32_t i = i0 + c;
            if (i < n) cnt += (i == 0 || k[i] != k[i - 1]) ? 1u : 0u;
        }
        uint32_t total;
        block_exclusive_sum<uint32_t>(cnt, sm.scan, total);
        if (threadIdx.x == 0) a.tile_heads[t] = total;
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) *sg.count = 0;
    if (!(alive = coop_barrier(a.bar, gen, sm.abort_flag, a.w.errflags))) return cur;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        if (threadIdx.x == 0) {
            uint32_t b = 0;
            for (int tt = 0; tt < t; ++tt) b += a.tile_heads[tt];
            sm.tile_base = b;
        }
        uint32_t cnt = 0;
        const uint32_t i0 = (uint32_t)t * kSortTile + threadIdx.x * kSortChunks;
#pragma unroll
        for (int c = 0; c < kSortChunks; ++c) {
            const uint32_t i = i0 + c;
            if (i < n) cnt += (i == 0 || k[i] != k[i - 1]) ? 1u : 0u;
        }
        uint32_t total;
        uint32_t s = block_exclusive_sum<uint32_t>(cnt, sm.scan, total) + sm.tile_base;  // (its barriers publish tile_base)
#pragma unroll
        for (int c = 0; c < kSortChunks; ++c) {
            const uint32_t i = i0 + c;
            if (i < n) {
                const bool head = (i == 0) || (k[i] != k[i - 1]);
                if (head) ++s;
                rle_fill_one(i, s - 1u, head, k, v, n, sg);
            }
        }
        __syncthreads();
    }
    alive = coop_barrier(a.bar, gen, sm.abort_flag, a.w.errflags);
    return cur;
}

__global__ void __launch_bounds__(kThreads) win_update_coop(const CoopArgs a) {
    __shared__ CoopSmem sm;
    if (threadIdx.x == 0) sm.abort_flag = 0;
    __syncthreads();
    unsigned gen = 0;
    bool alive = true;
    const uint32_t B = (uint32_t)a.B;
    const uint32_t n = a.directed ? 2u * B : 4u * B;
    const uint32_t gtid = blockIdx.x * kThreads + threadIdx.x, gsize = gridDim.x * kThreads;
    const bool outlists = a.directed && a.wo.vmeta != nullptr;
    if (gtid == 0 && outlists) *a.njobsB = 0;
    for (uint32_t i = gtid; i < B; i += gsize)
        batch_entries_one(i, a.log, a.W, a.log_start, a.arriving, a.B, a.directed, a.w.V, a.akey[0], a.aval[0], a.bkey[0],
                          a.bval[0], a.w.errflags, a.perm);
    {
        const uint32_t np = *a.npend_prev;
        for (uint32_t j = gtid; j < np; j += gsize) pool_reclaim_one(j, a.pend_prev, a.w.fr);
    }
    if (!coop_barrier(a.bar, gen, sm.abort_flag, a.w.errflags)) return;
    const int ra = coop_sort_and_rle(a, a.akey, a.aval, n, a.segA, sm, gen, alive);
    if (!alive) return;
    int rb = ra;
    if (a.directed) {
        rb = coop_sort_and_rle(a, a.bkey, a.bval, n, a.segB, sm, gen, alive);
        if (!alive) return;
    }
    // expire / reserve / grow per touched vertex; new out-degrees (different arrays: no barrier in between)
    const uint32_t nsegA = *a.segA.count, nsegB = *a.segB.count;
    for (uint32_t s = gtid; s < nsegA; s += gsize) plan_one(s, a.segA, a.w, a.ins_pos, a.jobs, a.njobs);
    for (uint32_t s = gtid; s < nsegB; s += gsize) out_degree_one(s, a.segB, a.w.outdeg, a.seg_d0);
    if (outlists)
        for (uint32_t s = gtid; s < nsegB; s += gsize) plan_one(s, a.segB, a.wo, a.ins_posB, a.jobsB, a.njobsB);
    if (!coop_barrier(a.bar, gen, sm.abort_flag, a.w.errflags)) return;
    const uint32_t gw = gtid >> 5, nw = gsize >> 5;
    for (int side = 0; side < (outlists ? 2 : 1); ++side) {
        const RelocJob *jobs = side ? a.jobsB : a.jobs;
        const uint32_t nj = side ? *a.njobsB : *a.njobs;
        for (uint32_t j = gw; j < nj; j += nw) {  // one warp per relocated ring
            const RelocJob jb = jobs[j];
            for (uint32_t k = threadIdx.x & 31; k < jb.len; k += 32)
                a.w.pool[jb.new_base + k] = a.w.pool[jb.old_base + ((jb.old_head + k) & (jb.old_cap - 1u))];
        }
    }
    const uint32_t *ka = a.akey[ra], *va = a.aval[ra];
    for (uint32_t i = gtid; i < n; i += gsize) insert_one(i, ka, va, a.segA, a.ins_pos, a.w);
    if (outlists) {
        const uint32_t *kb = a.bkey[rb], *vb = a.bval[rb];
        for (uint32_t i = gtid; i < n; i += gsize) insert_one(i, kb, vb, a.segB, a.ins_posB, a.wo);
    }
}

}  // namespace dppr
