// pull.cuh -- DENSE iterations of the level-synchronous push, executed in gather ("pull") form.
//
// One level-synchronous iteration of variant 0 (push.cuh) pops every frontier vertex u -- p[u] += a x, r[u] -= x for the
// popped amount x -- and adds (1-a) x / (outdeg(w)+1) to r[w] of every in-neighbour w: one random FP64 atomic per
// traversed edge.  When the frontier covers a large part of the graph the same iteration is cheaper the other way
// round: publish the popped amounts as a dense vector x (0 for vertices outside the frontier) and let every vertex w
// GATHER
//     r'[w] = r[w] + (1-a) / (outdeg(w)+1) * sum_{u in out(w)} x[u]
// from its OUT-list.  No atomics, no frontier queue, no owner search; each (w, source) is written by exactly one lane.
// The reference has no counterpart (it always scatters, gpu/ExpandRev.cuh).
//
// Round 2 rewrite (ncu on BASELINE configs[3], 125 sources: 29 % of DRAM peak, 5.2 G spill loads, barrier and
// long-scoreboard stalls; scripts/micro/gather.cu: random rows >= 256 B stream at 5-7 TB/s, anything narrower is bound
// by ~25 rows/ns):
//
//   * x holds bf16 values.  A push may move ANY part of a residual and keeps the invariant
//     (p + a r = a e_s + (1-a)/(d+1) sum p), so a sweep pops the residual TRUNCATED TOWARDS ZERO to bf16 and
//     leaves the remainder (< 2^-7 of it, same sign) in r, where it is popped by a later sweep if it is still above eps.
//     Every gathered value is exactly what was added to p / a, the sums are formed in FP64: the invariant stays exact to
//     rounding, the contraction per sweep goes from 0.85 to ~0.854, and the gather traffic -- the dominant term -- drops
//     4x (2 instead of 8 bytes per (edge, source)).  With one source the whole x vector of the Twitter-shaped window is
//     83 MB and stays L2-resident (126 MB) instead of being a 333 MB random-sector DRAM stream.
//   * p, r and x are VERTEX-major ([V][Sr]): the lanes that share a vertex read one contiguous row piece per out-list
//     entry (16 B = 8 sources per lane, up to 32 lanes = 512 B per request), and the per-vertex work (r row, p row, x
//     row) is coalesced too.
//   * no CTA-wide barrier in a sweep: a WARP owns 32 / G consecutive vertices (G = lanes per vertex).  Short lists are
//     walked by the owning lane group with several (PullUnroll) independent gathers in flight; lists of warp_min or more entries
//     are walked by the whole warp and reduced by shuffles; lists of big_min or more entries are cut into chunks that any
//     warp of the grid takes, partial sums meet in `bigacc` by FP64 atomics and the warp that completes the last chunk
//     finishes the vertex.  One grid barrier per sweep.
//   * the pop is still deferred by one sweep: sweep k decides x_next[w]; sweep k+1 adds a x to p[w].  Leaving dense mode
//     gives the undelivered x back to r and compacts those entries into an ordinary (un-popped) frontier queue.
//
// Included by push.cuh (needs PushArgs / PushSmem / the grid barrier).
#pragma once

namespace dppr {

#ifndef DPPR_PULL_UNROLL1
#define DPPR_PULL_UNROLL1 8
#endif
#ifndef DPPR_PULL_UNROLL8
#define DPPR_PULL_UNROLL8 4
#endif
// independent gathers in flight per lane: 2-byte entries with one source, 16-byte row pieces with several
template <int SB> struct PullUnroll { static constexpr int value = SB == 1 ? DPPR_PULL_UNROLL1 : DPPR_PULL_UNROLL8; };

// ---- bf16 pop amounts ----------------------------------------------------------------------------------------------
// truncation towards zero: |x| <= |r| and the remainder keeps the sign of r (a phase never creates residual of the
// other sign); values below the smallest normal float are not popped (they stay in r)
__device__ __forceinline__ uint32_t bf16_trunc(double r) {
    const uint32_t f = __float_as_uint(__double2float_rz(r)) >> 16;
    return (f & 0x7f80u) ? f : 0u;
}
__device__ __forceinline__ double bf16_value(uint32_t h) { return (double)__uint_as_float(h << 16); }

// x is gathered at random and re-read every sweep: keep it in L2 (evict_last), and let everything that merely streams
// through -- out-list slots, ring metadata -- leave first (ld.cs)
__device__ __forceinline__ unsigned long long l2_keep_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
#ifndef DPPR_PULL_HINTS
#define DPPR_PULL_HINTS 3   // bit 0: streaming (evict-first) loads of slots / metadata, bit 1: evict_last gathers of x
#endif
template <class T> __device__ __forceinline__ T pl_ldcs(const T *p) { return (DPPR_PULL_HINTS & 1) ? __ldcs(p) : __ldcg(p); }
// "this sector will be wanted soon": no register, no scoreboard -- the later load finds it in L2
#ifndef DPPR_PULL_PREFETCH
#define DPPR_PULL_PREFETCH 1
#endif
__device__ __forceinline__ void pl_prefetch_l2(const void *p) {
    if (DPPR_PULL_PREFETCH) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// a lane's piece of an x row: SB = 1 -> one bf16 (single source), SB = 8 -> 16 bytes
template <int SB> struct XPiece;
template <> struct XPiece<1> {
    uint32_t h;
    __device__ __forceinline__ bool any() const { return h != 0u; }
    __device__ __forceinline__ uint32_t get(int) const { return h; }
};
template <> struct XPiece<8> {
    uint4 v;
    __device__ __forceinline__ bool any() const { return (v.x | v.y | v.z | v.w) != 0u; }
    __device__ __forceinline__ uint32_t get(int j) const {
        const uint32_t w = j < 2 ? v.x : j < 4 ? v.y : j < 6 ? v.z : v.w;
        return (j & 1) ? (w >> 16) : (w & 0xffffu);
    }
};
template <int SB> __device__ __forceinline__ XPiece<SB> x_zero();
template <> __device__ __forceinline__ XPiece<1> x_zero<1>() { return XPiece<1>{0u}; }
template <> __device__ __forceinline__ XPiece<8> x_zero<8>() { return XPiece<8>{make_uint4(0u, 0u, 0u, 0u)}; }

// gathered at random (evict_last)
template <int SB> __device__ __forceinline__ XPiece<SB> x_gather(const uint16_t *x, size_t elem);
template <> __device__ __forceinline__ XPiece<1> x_gather<1>(const uint16_t *x, size_t elem) {
    unsigned short h;
    if (DPPR_PULL_HINTS & 2) asm volatile("ld.global.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(h) : "l"(x + elem), "l"(l2_keep_policy()));
    else h = x[elem];
    return XPiece<1>{(uint32_t)h};
}
template <> __device__ __forceinline__ XPiece<8> x_gather<8>(const uint16_t *x, size_t elem) {
    uint4 v;
    if (DPPR_PULL_HINTS & 2)
        asm volatile("ld.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(x + elem), "l"(l2_keep_policy()));
    else v = *reinterpret_cast<const uint4 *>(x + elem);
    return XPiece<8>{v};
}
// the vertex's own row piece (streamed)
template <int SB> __device__ __forceinline__ XPiece<SB> x_load(const uint16_t *x, size_t elem);
template <> __device__ __forceinline__ XPiece<1> x_load<1>(const uint16_t *x, size_t elem) { return XPiece<1>{(uint32_t)__ldcg(x + elem)}; }
template <> __device__ __forceinline__ XPiece<8> x_load<8>(const uint16_t *x, size_t elem) { return XPiece<8>{__ldcg(reinterpret_cast<const uint4 *>(x + elem))}; }
template <int SB> __device__ __forceinline__ void x_store(uint16_t *x, size_t elem, const uint32_t (&h)[SB]);
template <> __device__ __forceinline__ void x_store<1>(uint16_t *x, size_t elem, const uint32_t (&h)[1]) { __stcg(x + elem, (uint16_t)h[0]); }
template <> __device__ __forceinline__ void x_store<8>(uint16_t *x, size_t elem, const uint32_t (&h)[8]) {
    __stcg(reinterpret_cast<uint4 *>(x + elem), make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16)));
}

template <int SB>
__device__ __forceinline__ void x_accumulate(const XPiece<SB> &v, double (&acc)[SB], uint32_t &nz) {
#pragma unroll
    for (int j = 0; j < SB; ++j) {
        const uint32_t h = v.get(j);
        acc[j] += bf16_value(h);
        nz += h ? 1u : 0u;
    }
}

// ---- geometry ------------------------------------------------------------------------------------------------------
// A unit = (vertex w, SB sources starting at s0).  G = 2^gshift adjacent lanes share a vertex (their pieces are contiguous);
// a warp holds 32 / G consecutive vertices, a tile (the granularity of the active-tile list) kThreads / G of them; rows
// wider than SB * G sources are covered by several chunk groups, each with its own tiles.
struct PullGeom {
    uint32_t gs, G, vpw, vpt, nCG, tpc, ntiles;
};
template <int SB>
__device__ __forceinline__ PullGeom pull_geom(const PushArgs &a) {
    PullGeom q;
    q.gs = SB == 1 ? 0u : (uint32_t)a.pull_gshift;
    q.G = 1u << q.gs;
    q.vpw = 32u >> q.gs;
    q.vpt = (uint32_t)kThreads >> q.gs;
    q.nCG = SB == 1 ? 1u : ((uint32_t)a.Sr + SB * q.G - 1u) / (SB * q.G);
    q.tpc = ((uint32_t)a.V + q.vpt - 1u) / q.vpt;
    q.ntiles = q.tpc * q.nCG;
    return q;
}

constexpr uint32_t kBigDone = 0, kBigChunks = 1, kBigLen = 2;   // HubItem::pad slots used by the grid tier

// relaxation factor of an accelerated sweep: kept in shared memory (set by dense_body), not in registers -- the gather
// loops run at the register limit
__device__ __forceinline__ double sm_omega() {
    extern __shared__ __align__(16) unsigned char dppr_dyn_smem[];
    return *reinterpret_cast<const double *>(dppr_dyn_smem);
}

// A sweep pushes residuals of BOTH signs (the reference's two phases exist for the scatter form, whose threshold-crossing
// dedupe needs same-signed adds within a phase; a gather decides every (vertex, source) exactly once per sweep).  After a
// batch the repaired residuals have both signs and decay by the same linear process: one episode takes them down together
// instead of one ~40-sweep episode per phase.
__device__ __forceinline__ bool dense_legal(double x, double eps) { return fabs(x) > eps; }

// everything that happens once per unit: the deferred pop, the new residual, membership in the next frontier.
// Returns the number of sources of the unit that are in the next frontier.
template <int SB, bool ACCEL>
__device__ __forceinline__ uint32_t pull_finish_unit(const PushArgs &a, int phase, uint32_t w, uint32_t s0, uint32_t len,
                                                     const XPiece<SB> &xc, const double (&acc)[SB], uint16_t *xn,
                                                     unsigned long long &next_edges) {
    const size_t row = (size_t)w * (size_t)a.Sr + s0;
    uint32_t out[SB];
    bool touched = xc.any();
#pragma unroll
    for (int j = 0; j < SB; ++j) { out[j] = 0u; touched |= acc[j] != 0.0; }
    uint32_t legal = 0;
    if (touched) {
        const double scale = (1.0 - a.alpha) / (double)(len + 1u);
        if (xc.any()) {  // the pop decided by the previous sweep
            if constexpr (SB == 1) {
                a.p[row] += a.alpha * bf16_value(xc.get(0));
            } else {
                // Several sources: the popped amounts of an episode are summed per pair in FP32 (a.pacc, zero outside an
                // episode) and folded into p once, when the episode ends (pull_fold): 8 instead of 16 bytes of read-modify-write
                // per pair and sweep, on the stream that is two thirds of a full sweep's DRAM traffic.  The amounts are bf16
                // values, so an FP32 sum of them is almost always EXACT; TwoSum tells when it is not, and that amount goes
                // to p directly.  p therefore ends up with the same exact sum of amounts as before, times alpha.
                float pa[SB];
#pragma unroll
                for (int j = 0; j < SB; j += 4) {
                    const float4 v4 = *reinterpret_cast<const float4 *>(a.pacc + row + j);
                    pa[j] = v4.x; pa[j + 1] = v4.y; pa[j + 2] = v4.z; pa[j + 3] = v4.w;
                }
#pragma unroll
                for (int j = 0; j < SB; ++j) {
                    const float av = __uint_as_float(xc.get(j) << 16);
                    const float t = pa[j] + av;
                    const float bb = t - pa[j];
                    const float err = (pa[j] - (t - bb)) + (av - bb);
                    if (err == 0.f) pa[j] = t;
                    else a.p[row + j] += a.alpha * (double)av;  // (rare: the FP32 sum would have rounded)
                }
#pragma unroll
                for (int j = 0; j < SB; j += 4)
                    *reinterpret_cast<float4 *>(a.pacc + row + j) = make_float4(pa[j], pa[j + 1], pa[j + 2], pa[j + 3]);
            }
        }
        double rw[SB];
        if (SB == 1) {
            rw[0] = a.r[row];
        } else {
#pragma unroll
            for (int j = 0; j < SB; j += 2) {
                const double2 rv = *reinterpret_cast<const double2 *>(a.r + row + j);
                rw[j] = rv.x; rw[j + 1] = rv.y;
            }
        }
#pragma unroll
        for (int j = 0; j < SB; ++j) {
            rw[j] += acc[j] * scale;
            if (dense_legal(rw[j], a.eps)) {
                // Chebyshev semi-iteration (dense_body): push d_k = w r_k + (w - 1) d_{k-1}, d_{k-1} = what this unit popped in
                // this sweep.  w = 1 is the plain push of the whole residual.
                double want = rw[j];
                if (ACCEL) {
                    const double omega = sm_omega();
                    want = omega * rw[j] + (omega - 1.0) * bf16_value(xc.get(j));
                }
                const uint32_t h = bf16_trunc(want);
                if (h) {
                    out[j] = h;
                    rw[j] -= bf16_value(h);  // the remainder (of either sign) stays behind
                    ++legal;
                }
            }
        }
        if (SB == 1) {
            a.r[row] = rw[0];
        } else {
#pragma unroll
            for (int j = 0; j < SB; j += 2) *reinterpret_cast<double2 *>(a.r + row + j) = make_double2(rw[j], rw[j + 1]);
        }
    }
    x_store<SB>(xn, row, out);
    if (legal)  // what a scatter iteration would traverse for w: its in-degree, once per legal source
        next_edges += (unsigned long long)legal * (a.vmeta_out == a.vmeta ? len : __ldg(&a.vmeta[w]).z);
    return legal;
}

// CTA-wide sums of one count and one edge total per thread -> one atomic each
__device__ __forceinline__ void pull_count_flush(PushSmem &sm, uint32_t mine, unsigned int *cnt_out,
                                                 unsigned long long mine_edges = 0, unsigned long long *edges_out = nullptr) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mine += __shfl_xor_sync(kFull, mine, off);
        mine_edges += __shfl_xor_sync(kFull, mine_edges, off);
    }
    if (threadIdx.x == 0) { sm.pl_cnt = 0; sm.pl_edges = 0; }
    __syncthreads();
    if (lane_id() == 0 && mine) { atomicAdd(&sm.pl_cnt, mine); atomicAdd(&sm.pl_edges, mine_edges); }
    __syncthreads();
    if (threadIdx.x == 0 && sm.pl_cnt) {
        atomicAdd(cnt_out, sm.pl_cnt);
        if (edges_out) atomicAdd(edges_out, sm.pl_edges);
    }
}

// ---- entering dense mode ---------------------------------------------------------------------------------------------
// x[0][w] = the part of r[w] that the first sweep pops (where r[w] is legal), the rest stays in r; x[1] = 0.  Also lists
// the ACTIVE tiles -- a vertex without out-edges receives no adds, so a tile whose vertices have neither out-edges nor a
// legal residual stays all-zero in both x buffers for the whole episode and is never visited again (72 % of the
// Twitter-shaped window's vertices) -- and the grid tier: the (vertex, chunk group) pairs whose out-list is cut into chunks.
// Tiles are listed in a scrambled order: the heavy ones (heads of the relabel blocks) sit at a regular stride.
template <int SB>
__device__ __noinline__ void pull_build(const PushArgs &a, PushSmem &sm, PushCtrl *c, int phase, unsigned int *cnt_out) {
    const PullGeom q = pull_geom<SB>(a);
    const uint32_t V = (uint32_t)a.V;
    uint32_t legal = 0;
    unsigned long long ep_slots = 0, ep_pairs = 0, ep_units = 0;  // what ONE sweep over the active tiles moves
    for (uint32_t t = blockIdx.x; t < q.ntiles; t += gridDim.x) {
        const uint32_t tile = (uint32_t)(((unsigned long long)t * a.pull_tile_mul) % q.ntiles);
        const uint32_t cg = tile / q.tpc;
        const uint32_t w = (tile - cg * q.tpc) * q.vpt + (threadIdx.x >> q.gs), g = threadIdx.x & (q.G - 1u);
        const uint32_t s0 = (cg * q.G + g) * SB;
        const bool have = w < V && s0 < (uint32_t)a.Sr;
        bool active = false;
        uint32_t wlen = 0;
        if (have) {
            wlen = (uint32_t)__ldg(&a.outdeg[w]);
            active = wlen != 0;
            const size_t row = (size_t)w * (size_t)a.Sr + s0;
            uint32_t out[SB], zero[SB];
            double rw[SB];
            bool any = false;
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                out[j] = 0u; zero[j] = 0u;
                rw[j] = __ldcg(&a.r[row + j]);
                if (dense_legal(rw[j], a.eps)) {
                    const uint32_t h = bf16_trunc(rw[j]);
                    if (h) { out[j] = h; rw[j] -= bf16_value(h); ++legal; any = true; }
                }
            }
            if (any) {
                active = true;
#pragma unroll
                for (int j = 0; j < SB; ++j)
                    if (out[j]) __stcg(&a.r[row + j], rw[j]);
            }
            x_store<SB>(a.x[0], row, out);
            x_store<SB>(a.x[1], row, zero);
            if (g == 0 && wlen >= (uint32_t)a.pull_big_min) {  // grid tier: one entry per (vertex, chunk group)
                const uint32_t nch = (wlen + (uint32_t)a.pull_big_chunk - 1u) / (uint32_t)a.pull_big_chunk;
                const unsigned long long old = atomicAdd(&c->bigpk, (1ull << 32) | nch);
                const uint32_t hp = (uint32_t)(old >> 32);
                if (hp < a.bigcap) {
                    a.big[hp].item = ((unsigned long long)cg << 32) | w;
                    a.big[hp].chunk0 = (uint32_t)old;
                    a.big[hp].pad[kBigDone] = 0u;
                    a.big[hp].pad[kBigChunks] = nch;
                    a.big[hp].pad[kBigLen] = wlen;
                } else {
                    atomicOr(&c->errflags, kErrHubQ);
                }
            }
        }
        // the tile's weight (out-list entries its vertices hold) decides its class: heavy tiles are handed out first
        {
            uint32_t wsum = (have && g == 0) ? wlen : 0u;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) wsum += __shfl_xor_sync(kFull, wsum, off);
            const bool wact = __any_sync(kFull, active);
            if (threadIdx.x == 0) { sm.pl_cnt = 0; sm.pl_n = 0; }
            __syncthreads();
            if (lane_id() == 0) {
                if (wsum) atomicAdd(&sm.pl_cnt, wsum);
                if (wact) atomicOr(&sm.pl_n, 1u);
            }
            __syncthreads();
        }
        const bool tile_active = sm.pl_n != 0;
        if (tile_active) {
            if (threadIdx.x == 0) {
                const double mean = (double)a.avg_indeg * (double)q.vpt;  // (entries per tile if all lists were equal)
                const int cls = (double)sm.pl_cnt >= 4.0 * mean ? 0 : (double)sm.pl_cnt >= 1.5 * mean ? 1 : 2;
                a.tile_list[(size_t)cls * a.tile_list_cap + atomicAdd(&c->ntiles_b[cls], 1u)] = tile;
                atomicAdd(&c->ntiles_active, 1u);
            }
            if (have) {
                const uint32_t nreal = s0 < (uint32_t)a.S ? min((uint32_t)SB, (uint32_t)a.S - s0) : 0u;
                ep_units += nreal;
                ep_pairs += (unsigned long long)wlen * nreal;
                if (g == 0) ep_slots += wlen;
            }
        }
        __syncthreads();  // (sm.pl_* are reused by the next tile)
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ep_slots += __shfl_xor_sync(kFull, ep_slots, off);
        ep_pairs += __shfl_xor_sync(kFull, ep_pairs, off);
        ep_units += __shfl_xor_sync(kFull, ep_units, off);
    }
    if (lane_id() == 0 && ep_units) {
        atomicAdd(&c->ep_slots, ep_slots);
        atomicAdd(&c->ep_pairs, ep_pairs);
        atomicAdd(&c->ep_units, ep_units);
    }
    pull_count_flush(sm, legal, cnt_out);
}

// ---- list walks ------------------------------------------------------------------------------------------------------
// entries first, first + step, ... < last of the ring {base, head, mask}: each lane group gathers ITS piece (column c0 of the
// row) of x[slot], kPullUnroll slots in flight
template <int SB>
__device__ __forceinline__ void pull_walk(const PushArgs &a, const uint16_t *xcur, uint32_t base, uint32_t head, uint32_t mask,
                                          uint32_t first, uint32_t last, uint32_t step, uint32_t c0, double (&acc)[SB], uint32_t &nz) {
    // software-pipelined: the slots of round i + 1 are requested together with the row pieces of round i, so a round
    // costs one memory round trip, not two (slot -> row piece)
    constexpr int U = PullUnroll<SB>::value;
    if (first >= last) return;
    uint32_t u[U];
#pragma unroll
    for (int i = 0; i < U; ++i) u[i] = first + i * step < last ? (uint32_t)pl_ldcs(&a.pool[base + ((head + first + i * step) & mask)]) : 0xffffffffu;
    for (uint32_t k = first; k < last; k += U * step) {
        XPiece<SB> v[U];
        uint32_t un[U];
#pragma unroll
        for (int i = 0; i < U; ++i) v[i] = u[i] != 0xffffffffu ? x_gather<SB>(xcur, (size_t)u[i] * (size_t)a.Sr + c0) : x_zero<SB>();
        const uint32_t kn = k + U * step;
#pragma unroll
        for (int i = 0; i < U; ++i) un[i] = kn + i * step < last ? (uint32_t)pl_ldcs(&a.pool[base + ((head + kn + i * step) & mask)]) : 0xffffffffu;
#pragma unroll
        for (int i = 0; i < U; ++i) x_accumulate<SB>(v[i], acc, nz);
#pragma unroll
        for (int i = 0; i < U; ++i) u[i] = un[i];
    }
}

// sum over the lane groups of a warp; afterwards every lane holds the total of its column
template <int SB>
__device__ __forceinline__ void pull_reduce_groups(double (&part)[SB], uint32_t G) {
#pragma unroll
    for (int j = 0; j < SB; ++j)
        for (uint32_t off = 16; off >= G; off >>= 1) part[j] += __shfl_xor_sync(kFull, part[j], off);
}

// ---- one sweep -------------------------------------------------------------------------------------------------------
// Work items -- first the chunks of the grid tier (the longest tasks), then the active tiles, heavy classes first -- are
// handed out one at a time by an atomic counter (round 2: with the static tile -> CTA assignment 20-35 % of a sweep was
// spent waiting at its closing barrier); the index of the next item is requested before the current one is processed, so
// the atomic's round trip is hidden.
// `gath` counts the gathered x entries that were non-zero: exactly the (edge, source) pairs the push form would have
// traversed.
template <int SB>
__device__ __forceinline__ uint32_t pull_tile_at(const PushArgs &a, uint32_t j, uint32_t n0, uint32_t n1) {
    if (j < n0) return __ldcg(&a.tile_list[j]);
    j -= n0;
    if (j < n1) return __ldcg(&a.tile_list[(size_t)a.tile_list_cap + j]);
    return __ldcg(&a.tile_list[2 * (size_t)a.tile_list_cap + (j - n1)]);
}

// per-warp running totals of a sweep
struct PullAcc {
    uint32_t legal = 0, nz = 0;
    unsigned long long next_edges = 0;
};

// one chunk of a long out-list, by one warp; the warp that completes the last chunk of a vertex finishes it
template <int SB, bool ACCEL>
__device__ __forceinline__ void pull_do_chunk(const PushArgs &a, const PullGeom &q, int phase, const uint16_t *xcur, uint16_t *xnext,
                                              uint32_t cidx, uint32_t nh, PullAcc &t) {
    const uint32_t lane = lane_id(), grp = lane >> q.gs, g = lane & (q.G - 1u);
    uint32_t lo = 0, hi = nh;  // last list entry with chunk0 <= cidx
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldcg(&a.big[mid].chunk0) <= cidx) lo = mid; else hi = mid;
    }
    const unsigned long long item = __ldcg(&a.big[lo].item);
    const uint32_t w = (uint32_t)item, cg = (uint32_t)(item >> 32);
    const uint32_t s0 = (cg * q.G + g) * SB;
    const uint4 m = __ldg(&a.vmeta_out[w]);
    const uint32_t e0 = (cidx - __ldcg(&a.big[lo].chunk0)) * (uint32_t)a.pull_big_chunk;
    const uint32_t e1 = min(m.z, e0 + (uint32_t)a.pull_big_chunk);
    double part[SB];
#pragma unroll
    for (int jj = 0; jj < SB; ++jj) part[jj] = 0.0;
    if (s0 < (uint32_t)a.Sr) pull_walk<SB>(a, xcur, m.x, m.y, m.w - 1u, e0 + grp, e1, q.vpw, s0, part, t.nz);
    pull_reduce_groups<SB>(part, q.G);
    double *accrow = a.bigacc + (size_t)lo * (size_t)(q.G * SB);
    if (grp == 0 && s0 < (uint32_t)a.Sr) {
#pragma unroll
        for (int jj = 0; jj < SB; ++jj)
            if (part[jj] != 0.0) atomicAdd(&accrow[g * SB + jj], part[jj]);
    }
    __threadfence();  // the partial sums are out before the chunk is counted
    __syncwarp();
    uint32_t done = 0;
    if (lane == 0) done = atomicAdd(&a.big[lo].pad[kBigDone], 1u) + 1u;
    done = __shfl_sync(kFull, done, 0);
    if (done == __ldcg(&a.big[lo].pad[kBigChunks])) {
        __threadfence();
        if (grp == 0 && s0 < (uint32_t)a.Sr) {
            double acc[SB];
#pragma unroll
            for (int jj = 0; jj < SB; ++jj) {
                acc[jj] = __ldcg(&accrow[g * SB + jj]);
                __stcg(&accrow[g * SB + jj], 0.0);
            }
            const XPiece<SB> xc = x_load<SB>(xcur, (size_t)w * (size_t)a.Sr + s0);
            t.legal += pull_finish_unit<SB, ACCEL>(a, phase, w, s0, m.z, xc, acc, xnext, t.next_edges);
        }
        if (lane == 0) a.big[lo].pad[kBigDone] = 0u;  // (ready for the next sweep)
    }
}

// the 32 / G consecutive vertices starting at wfirst (chunk group cg), by one warp
template <int SB, bool ACCEL>
__device__ __forceinline__ void pull_do_vertices(const PushArgs &a, const PullGeom &q, int phase, const uint16_t *xcur, uint16_t *xnext,
                                                 uint32_t wfirst, uint32_t cg, PullAcc &t) {
    const uint32_t V = (uint32_t)a.V;
    const uint32_t lane = lane_id(), grp = lane >> q.gs, g = lane & (q.G - 1u);
    const uint32_t w = wfirst + grp, s0 = (cg * q.G + g) * SB;
    const bool have = w < V && s0 < (uint32_t)a.Sr;
    XPiece<SB> xc = x_zero<SB>();
    double acc[SB];
#pragma unroll
    for (int jj = 0; jj < SB; ++jj) acc[jj] = 0.0;
    uint32_t len = 0, base = 0, head = 0, mask = 0;
    if (w < V) {
        const uint4 m = pl_ldcs(&a.vmeta_out[w]);  // (the length of an out-list IS the out-degree)
        base = m.x; head = m.y; len = m.z; mask = m.w - 1u;
    }
    if (have) xc = x_load<SB>(xcur, (size_t)w * (size_t)a.Sr + s0);
    if (SB > 1 && have) {
        // the unit's own r piece (64 bytes) is read when its list has been walked, a dependent DRAM round trip later: ask
        // for it now (pull_finish_unit's share of the stall samples was 25 %, profiles/README.md)
        const double *rp = a.r + (size_t)w * (size_t)a.Sr + s0;
        pl_prefetch_l2(rp); pl_prefetch_l2(rp + 4);
    }
    const int tier = len >= (uint32_t)a.pull_big_min ? 2 : (len >= (uint32_t)a.pull_warp_min && q.vpw > 1u) ? 1 : 0;
    if (SB > 1 && have && xc.any())  // ... and its piece of the popped-amount sums, which only a unit popped by the previous sweep touches
        pl_prefetch_l2(a.pacc + (size_t)w * (size_t)a.Sr + s0);
    if (tier == 0 && have && len) pull_walk<SB>(a, xcur, base, head, mask, 0u, len, 1u, s0, acc, t.nz);
    // lists of warp_min or more entries: the whole warp walks them, one after the other
    unsigned m1 = __ballot_sync(kFull, tier == 1 && g == 0 && w < V);
    while (m1) {
        const int L = __ffs(m1) - 1;
        m1 &= m1 - 1u;
        const uint32_t eb = __shfl_sync(kFull, base, L), eh = __shfl_sync(kFull, head, L), el = __shfl_sync(kFull, len, L),
                       em = __shfl_sync(kFull, mask, L);
        double part[SB];
#pragma unroll
        for (int jj = 0; jj < SB; ++jj) part[jj] = 0.0;
        if (s0 < (uint32_t)a.Sr) pull_walk<SB>(a, xcur, eb, eh, em, grp, el, q.vpw, s0, part, t.nz);
        pull_reduce_groups<SB>(part, q.G);
        if (grp == ((uint32_t)L >> q.gs)) {
#pragma unroll
            for (int jj = 0; jj < SB; ++jj) acc[jj] = part[jj];
        }
    }
    if (tier != 2 && have) t.legal += pull_finish_unit<SB, ACCEL>(a, phase, w, s0, len, xc, acc, xnext, t.next_edges);
}

template <int SB, bool ACCEL>
__device__ __forceinline__ void pull_sweep(const PushArgs &a, PushSmem &sm, PushCtrl *c, int phase, const uint16_t *xcur, uint16_t *xnext,
                           unsigned int *cnt_out, unsigned long long *edges_out, unsigned long long &gath, uint32_t sweep_index) {
    const PullGeom q = pull_geom<SB>(a);
    const uint32_t V = (uint32_t)a.V;
    const uint32_t lane = lane_id();
    PullAcc t;
    const unsigned long long bp = __ldcg(&c->bigpk);
    const uint32_t nh = min((uint32_t)(bp >> 32), a.bigcap), nchunks = nh ? (uint32_t)bp : 0u;
    const uint32_t n0 = __ldcg(&c->ntiles_b[0]), n1 = __ldcg(&c->ntiles_b[1]), n2 = __ldcg(&c->ntiles_b[2]);
    unsigned int *next = &c->work_next[sweep_index & 1u];
    if (blockIdx.x == 0 && threadIdx.x == 0) c->work_next[(sweep_index + 1u) & 1u] = 0u;  // (idle during this sweep)

    if (SB == 1) {
        // ---- one source: items handed to WARPS -- a warp takes a chunk, or a whole tile (256 vertices, 32 at a time) ----
        const uint32_t nwork = nchunks + n0 + n1 + n2;
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(next, 1u);
        j = __shfl_sync(kFull, j, 0);
        while (j < nwork) {
            uint32_t jn = 0;
            if (lane == 0) jn = atomicAdd(next, 1u);  // consumed at the bottom of the loop
            if (j < nchunks) {
                pull_do_chunk<SB, ACCEL>(a, q, phase, xcur, xnext, j, nh, t);
            } else {
                const uint32_t tile = pull_tile_at<SB>(a, j - nchunks, n0, n1);
                const uint32_t cg = tile / q.tpc;
                const uint32_t w0 = (tile - cg * q.tpc) * q.vpt;
                for (uint32_t sub = 0; sub < (uint32_t)kWarps && w0 + sub * q.vpw < V; ++sub)
                    pull_do_vertices<SB, ACCEL>(a, q, phase, xcur, xnext, w0 + sub * q.vpw, cg, t);
            }
            j = __shfl_sync(kFull, jn, 0);
        }
    } else if (a.pull_warp_units > 0) {
        // ---- several sources, items handed to WARPS (the default): a chunk of the grid tier, or `wu` consecutive warp-slices
        // (32 / G vertices each; 8 = one tile).  No CTA barrier inside the sweep: with items handed to CTAs (below) the warps of
        // a CTA spent 2.75 of 21.5 stall cycles per issued instruction waiting for the slowest warp of each item
        // (profiles/ncu_c4_s125_r02c.txt).  BASELINE configs[3], 125 sources: 120.5 -> 108-111 ms per batch with 8 / 16 / 32
        // slices per item (1: 117, 2: 113, 4: 111, 64: 116); configs[4], 8 sources: 163 -> 131 ms ----
        const uint32_t wu = (uint32_t)a.pull_warp_units;
        const uint32_t ntl = n0 + n1 + n2;
        const uint32_t nslices = ntl * (uint32_t)kWarps;  // (a tile = kWarps warp-slices)
        const uint32_t nwork = nchunks + (nslices + wu - 1) / wu;
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(next, 1u);
        j = __shfl_sync(kFull, j, 0);
        while (j < nwork) {
            uint32_t jn = 0;
            if (lane == 0) jn = atomicAdd(next, 1u);  // consumed at the bottom of the loop
            if (j < nchunks) {
                pull_do_chunk<SB, ACCEL>(a, q, phase, xcur, xnext, j, nh, t);
            } else {
                const uint32_t s_lo = (j - nchunks) * wu, s_hi = min(nslices, s_lo + wu);
                uint32_t tt_cur = 0xffffffffu, cg = 0, w0 = 0;
                for (uint32_t sl = s_lo; sl < s_hi; ++sl) {
                    const uint32_t tt = sl / (uint32_t)kWarps;
                    if (tt != tt_cur) {
                        tt_cur = tt;
                        const uint32_t tile = pull_tile_at<SB>(a, tt, n0, n1);
                        cg = tile / q.tpc;
                        w0 = (tile - cg * q.tpc) * q.vpt;
                    }
                    const uint32_t wf = w0 + (sl - tt * (uint32_t)kWarps) * q.vpw;
                    if (wf < V) pull_do_vertices<SB, ACCEL>(a, q, phase, xcur, xnext, wf, cg, t);
                }
            }
            j = __shfl_sync(kFull, jn, 0);
        }
        __syncthreads();
    } else {
        // ---- several sources, items handed to CTAs (tuning.pull_warp_units = -1; kept for A/B runs) -- kWarps chunks of the grid
        // tier, or tiles whose vertices the CTA's warps share (32 / G each).  (Early in round 2, before the pipelined walk and
        // the 3-CTA build, this beat whole tiles per warp 498 to 683 ms; re-measured at the end of the round it loses, see above.) ----
        const uint32_t tpi = max(1u, 32u / q.vpt);  // tiles per item: at least 32 vertices
        const uint32_t ntl = n0 + n1 + n2;
        const uint32_t ngroups = (nchunks + kWarps - 1) / kWarps;
        const uint32_t nwork = ngroups + (ntl + tpi - 1) / tpi;
        if (threadIdx.x == 0) sm.pl_n = atomicAdd(next, 1u);
        __syncthreads();
        uint32_t j = sm.pl_n;
        uint32_t flip = 0;
        while (j < nwork) {
            unsigned int *nslot = flip ? &sm.pl_n : &sm.gbase;  // (two words used alternately: one barrier per item)
            if (threadIdx.x == 0) *nslot = atomicAdd(next, 1u);
            if (j < ngroups) {
                const uint32_t cidx = j * kWarps + warp_id();
                if (cidx < nchunks) pull_do_chunk<SB, ACCEL>(a, q, phase, xcur, xnext, cidx, nh, t);
            } else {
                for (uint32_t tt = (j - ngroups) * tpi; tt < min(ntl, (j - ngroups + 1) * tpi); ++tt) {
                    const uint32_t tile = pull_tile_at<SB>(a, tt, n0, n1);
                    const uint32_t cg = tile / q.tpc;
                    const uint32_t wf = (tile - cg * q.tpc) * q.vpt + warp_id() * q.vpw;
                    if (wf < V) pull_do_vertices<SB, ACCEL>(a, q, phase, xcur, xnext, wf, cg, t);
                }
            }
            __syncthreads();
            j = *nslot;
            flip ^= 1u;
        }
        __syncthreads();
    }
    gath += t.nz;
    pull_count_flush(sm, t.legal, cnt_out, t.next_edges, edges_out);
}

// Out-of-line copies: a sweep inlined twice (plain and accelerated) into the episode loop shares ONE register allocation
// with it, and ptxas then spills inside the gather loops (measured: 8.4 instead of 5.6 ms per sweep on BASELINE
// configs[3]).  Only the single-source plain sweep stays inline (its parameters then stay in the constant bank: -14 %).
template <int SB, bool ACCEL>
__device__ __noinline__ void pull_sweep_out(const PushArgs &a, PushSmem &sm, PushCtrl *c, int phase, const uint16_t *xcur, uint16_t *xnext,
                                            unsigned int *cnt_out, unsigned long long *edges_out, unsigned long long &gath, uint32_t sweep_index) {
    pull_sweep<SB, ACCEL>(a, sm, c, phase, xcur, xnext, cnt_out, edges_out, gath, sweep_index);
}

// ---- end of an episode: p += alpha * (sum of the amounts the episode popped), sums back to zero (several sources) ------
template <int SB>
__device__ __noinline__ void pull_fold(const PushArgs &a, PushCtrl *c) {
    const PullGeom q = pull_geom<SB>(a);
    const uint32_t V = (uint32_t)a.V;
    const uint32_t n0 = __ldcg(&c->ntiles_b[0]), n1 = __ldcg(&c->ntiles_b[1]), ntiles = n0 + n1 + __ldcg(&c->ntiles_b[2]);
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint32_t tile = pull_tile_at<SB>(a, t, n0, n1);
        const uint32_t cg = tile / q.tpc;
        const uint32_t w = (tile - cg * q.tpc) * q.vpt + (threadIdx.x >> q.gs), g = threadIdx.x & (q.G - 1u);
        const uint32_t s0 = (cg * q.G + g) * SB;
        if (!(w < V && s0 < (uint32_t)a.Sr)) continue;
        const size_t row = (size_t)w * (size_t)a.Sr + s0;
#pragma unroll
        for (int j = 0; j < SB; j += 4) {
            const float4 v4 = *reinterpret_cast<const float4 *>(a.pacc + row + j);
            if (v4.x == 0.f && v4.y == 0.f && v4.z == 0.f && v4.w == 0.f) continue;
            double2 p0 = *reinterpret_cast<double2 *>(a.p + row + j), p1 = *reinterpret_cast<double2 *>(a.p + row + j + 2);
            p0.x += a.alpha * (double)v4.x; p0.y += a.alpha * (double)v4.y;
            p1.x += a.alpha * (double)v4.z; p1.y += a.alpha * (double)v4.w;
            *reinterpret_cast<double2 *>(a.p + row + j) = p0;
            *reinterpret_cast<double2 *>(a.p + row + j + 2) = p1;
            *reinterpret_cast<float4 *>(a.pacc + row + j) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}
template <>
__device__ __noinline__ void pull_fold<1>(const PushArgs &, PushCtrl *) {}  // (one source updates p in every sweep)

// ---- leaving dense mode ----------------------------------------------------------------------------------------------
// the non-zero entries of x are pops that were decided but not performed: give them back to r; those (source, vertex)
// pairs whose residual has the sign of the running phase are the (un-popped) frontier of the next scatter iteration, the
// others wait in qalt for the next phase
template <int SB>
__device__ __noinline__ void pull_compact(const PushArgs &a, PushSmem &sm, PushCtrl *c, int phase, const uint16_t *x, unsigned long long *qout,
                             unsigned int *cnt_out) {
    const PullGeom q = pull_geom<SB>(a);
    const uint32_t V = (uint32_t)a.V;
    const uint32_t n0 = __ldcg(&c->ntiles_b[0]), n1 = __ldcg(&c->ntiles_b[1]), ntiles = n0 + n1 + __ldcg(&c->ntiles_b[2]);
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint32_t tile = pull_tile_at<SB>(a, t, n0, n1);
        const uint32_t cg = tile / q.tpc;
        const uint32_t w = (tile - cg * q.tpc) * q.vpt + (threadIdx.x >> q.gs), g = threadIdx.x & (q.G - 1u);
        const uint32_t s0 = (cg * q.G + g) * SB;
        const bool have = w < V && s0 < (uint32_t)a.Sr;
        const size_t row = (size_t)w * (size_t)a.Sr + s0;
        XPiece<SB> xc = x_zero<SB>();
        if (have) xc = x_load<SB>(x, row);
#pragma unroll
        for (int j = 0; j < SB; ++j) {
            const uint32_t h = xc.get(j);
            bool mine = false;
            if (h) {
                const double rw = a.r[row + j] + bf16_value(h);
                a.r[row + j] = rw;
                mine = legal_push(rw, phase, a.eps);
                if (!mine && dense_legal(rw, a.eps)) {  // the other sign: a seed of the next phase (phase is 0 or 1 here)
                    const unsigned pos = atomicAdd(&c->nalt[(phase ^ 1) & 1], 1u);
                    if (pos < a.qcap) a.qalt[pos] = ((unsigned long long)(s0 + j) << 32) | w;
                    else atomicOr(&a.ctrl->errflags, kErrQueue);
                }
            }
            stage_push(mine, ((unsigned long long)(s0 + j) << 32) | w, sm, qout, cnt_out, a.qcap, a.ctrl);
        }
        stage_flush(sm, qout, cnt_out, a.qcap, a.ctrl);
    }
}

// The dense episode that replaces scatter iteration `it`.  On return the frontier of iteration it+1 is in
// q[(it+1)&1] / cnt[(it+1)%3] and no hubs are pending.
template <int SB>
__device__ __forceinline__ bool dense_body(const PushArgs &a, PushSmem &sm, PushCtrl *c, GridBar &gen, int phase, uint32_t it,
                                           unsigned long long hpk, unsigned long long &edges_acc, unsigned long long &gath,
                                           unsigned long long &pops_acc, uint32_t &iters_done, uint32_t &sweeps_done,
                                           float rate) {
    // hubs popped in iteration it-1 still owe their adds ...
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c->dcnt[0] = 0; c->dcnt[1] = 0; c->dcnt[2] = 0;
        c->dedges[0] = 0; c->dedges[1] = 0; c->dedges[2] = 0;
        c->bigpk = 0;
        c->ntiles_active = 0;
        c->ntiles_b[0] = 0; c->ntiles_b[1] = 0; c->ntiles_b[2] = 0;
        c->work_next[0] = 0; c->work_next[1] = 0;
        c->ep_slots = 0; c->ep_pairs = 0; c->ep_units = 0;
        c->nalt[(phase ^ 1) & 1] = 0;  // (this episode re-absorbs whatever an earlier one of this phase left for the next phase)
    }
    // ... and are simply UN-popped instead of being scattered edge by edge: r[u] += ru, p[u] -= a ru (one thread per hub;
    // r[u] may already hold adds of this iteration, hence the atomic).  The build pass below then finds u legal again and
    // the first sweep pushes it in gather form.  With 125 sources on the Orkut-shaped window the seeds of a batch are
    // mostly hubs: scattering them cost 60 ms per batch.
    {
        const HubItem *hin = a.hub[(it + 1) & 1];
        const uint32_t nh = min((uint32_t)(hpk >> 32), a.hcap);
        for (uint32_t h = blockIdx.x * kThreads + threadIdx.x; h < nh; h += gridDim.x * kThreads) {
            const unsigned long long item = __ldcg(&hin[h].item);
            const double ru = __ldcg(&hin[h].ru);
            const size_t idx = (size_t)(uint32_t)item * (size_t)a.Sr + (item >> 32);
            atomicAdd(&a.r[idx], ru);
            atomicAdd(&a.p[idx], -a.alpha * ru);
        }
        (void)edges_acc;
        if (blockIdx.x == 0 && threadIdx.x == 0) pops_acc -= nh;  // they are popped again by the first sweep
    }
    if (!grid_barrier(c, gen, sm)) return false;
    if (blockIdx.x == 0 && threadIdx.x == 0) c->cnt[(it + 1) % 3] = 0;  // the queues are rebuilt from x on the way out
    pull_build<SB>(a, sm, c, phase, &c->dcnt[0]);
    if (!grid_barrier(c, gen, sm)) return false;
    uint32_t k = 0;
    int cur = 0;
    // leave when scattering the next frontier is expected to be cheaper than one more sweep (measured costs if this
    // engine has them, the host's static estimate otherwise)
    const float sw_known = __ldcg(&c->sweep_ns);
    const unsigned long long t_ep0 = (blockIdx.x == 0 && threadIdx.x == 0) ? global_ns() : 0ull;
    // Chebyshev acceleration.  While (almost) every (vertex, source) pushes, a sweep is the linear iteration r <- M r with
    // M = (1-a) (D+I)^-1 A: ~25 sweeps of a plateau on which every residual just shrinks by 0.85.  A push of ANY amount keeps
    // the invariant, so the sweeps may follow the Chebyshev three-term recurrence instead -- the amount pushed is
    // d_k = w_k r_k + (w_k - 1) d_{k-1} -- which needs M's spectrum to be real: true for an undirected window, where M is
    // similar to a symmetric matrix and its eigenvalues lie in (-(1-a), 1-a).  Asymptotic factor per sweep 0.557 instead
    // of 0.85.  Every CTA derives w_k from the same frontier counts; off (w = 1) while the frontier is below accel_frac of
    // all pairs, where thresholding makes the iteration non-linear, and on directed graphs.
    const double rho2 = (1.0 - a.alpha) * (1.0 - a.alpha);
    const double accel_min = a.accel_frac > 0.0 ? a.accel_frac * (double)a.V * (double)a.S : 1e300;
    double omega = 1.0;
    uint32_t kacc = 0;
    while (true) {
        const uint32_t n = __ldcg(&c->dcnt[k % 3]);
        if (n == 0) break;
        if ((double)n >= accel_min) {
            ++kacc;
            omega = kacc == 1 ? 1.0 : kacc == 2 ? 1.0 / (1.0 - 0.5 * rho2) : 1.0 / (1.0 - 0.25 * rho2 * omega);
        } else {
            kacc = 0;
            omega = 1.0;
        }
        {
            extern __shared__ __align__(16) unsigned char dppr_dyn_smem[];
            __syncthreads();
            if (threadIdx.x == 0) *reinterpret_cast<double *>(dppr_dyn_smem) = omega;
            __syncthreads();
        }
        if (k > 0) {
            const unsigned long long ne = __ldcg(&c->dedges[k % 3]);
            const bool fits = (double)n < 0.5 * (double)a.qcap;  // (the frontier must fit the queue it is compacted into)
            const bool leave = fits && ((rate > 0.f && sw_known > 0.f) ? (double)ne * (double)rate < 0.8 * (double)sw_known
                                                                        : ne < a.dense_exit_edges);
            if (leave) break;
        }
        if ((int)iters_done >= a.max_iters || k >= 4096u) {  // (an episode needs a few dozen sweeps)
            if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&c->errflags, kErrWatchdog);
            return false;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            c->dcnt[(k + 2) % 3] = 0;
            c->dedges[(k + 2) % 3] = 0;
            pops_acc += n;
            c->dpops += n;
            if (a.iterlog && (int)iters_done < a.iterlog_cap) {
                const unsigned long long t = global_ns();
                a.iterlog[iters_done] = make_uint4(n, 0xffffffffu, (uint32_t)t, (uint32_t)(t >> 32));
            }
        }
        if (omega != 1.0) pull_sweep_out<SB, true>(a, sm, c, phase, a.x[cur], a.x[cur ^ 1], &c->dcnt[(k + 1) % 3], &c->dedges[(k + 1) % 3], gath, k);
        else if (SB == 1) pull_sweep<SB, false>(a, sm, c, phase, a.x[cur], a.x[cur ^ 1], &c->dcnt[(k + 1) % 3], &c->dedges[(k + 1) % 3], gath, k);
        else pull_sweep_out<SB, false>(a, sm, c, phase, a.x[cur], a.x[cur ^ 1], &c->dcnt[(k + 1) % 3], &c->dedges[(k + 1) % 3], gath, k);
        if (!grid_barrier(c, gen, sm)) return false;
        cur ^= 1;
        ++k;
        ++iters_done;
        ++sweeps_done;
    }
    if (k > 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        c->sweep_ns = (float)((double)(global_ns() - t_ep0) / (double)k);
        c->walk_slots += (unsigned long long)k * __ldcg(&c->ep_slots);  // (written before the barrier that followed pull_build)
        c->walk_pairs += (unsigned long long)k * __ldcg(&c->ep_pairs);
        c->units += (unsigned long long)k * __ldcg(&c->ep_units);
    }
    if (SB > 1 && k > 0) pull_fold<SB>(a, c);  // (p, pacc: nothing the compaction below reads or writes)
    if (__ldcg(&c->dcnt[k % 3]) != 0)
        pull_compact<SB>(a, sm, c, phase, a.x[cur], a.q[(it + 1) & 1], &c->cnt[(it + 1) % 3]);
    return grid_barrier(c, gen, sm);
}

// (not inlined, and its in/out scalars travel in one struct: the scatter path keeps its register allocation)
struct DenseIO {
    unsigned long long edges_acc, gath, pops_acc;
    uint32_t iters_done, sweeps_done;
    GridBar gen;
    float rate;  // scatter cost per edge the caller decided with (0 = not measured yet)
};
// How the argument block reaches the out-of-line episode (measured, profiles/README.md): several sources -- by REFERENCE to
// the kernel's __grid_constant__ parameter (no 400-byte stack copy whose fields the gather loops re-read as local loads:
// BASELINE configs[3], 125 sources, 108.5 -> 99.5 ms per batch); one source -- by VALUE (its sweep is inlined and keeps
// what it needs in registers; through the reference it reads the parameter window with generic loads: 37 -> 43 ms on
// configs[4])
template <int SB> struct DenseArgPass { using type = const PushArgs &; };
template <> struct DenseArgPass<1> { using type = const PushArgs; };
template <int SB>
__device__ __noinline__ bool dense_mode(typename DenseArgPass<SB>::type a, PushSmem &sm, PushCtrl *c, int phase, uint32_t it,
                                        unsigned long long hpk, DenseIO &io) {  // `a` BY VALUE: a reference would force
    // the caller to keep its kernel parameters in local memory instead of the constant bank, for the scatter path too
    GridBar gen = io.gen;
    unsigned long long edges_acc = io.edges_acc, gath = io.gath, pops_acc = io.pops_acc;
    uint32_t iters_done = io.iters_done, sweeps_done = io.sweeps_done;
    const bool ok = dense_body<SB>(a, sm, c, gen, phase, it, hpk, edges_acc, gath, pops_acc, iters_done, sweeps_done, io.rate);
    io.gen = gen; io.edges_acc = edges_acc; io.gath = gath; io.pops_acc = pops_acc;
    io.iters_done = iters_done; io.sweeps_done = sweeps_done;
    return ok;
}

}  // namespace dppr
