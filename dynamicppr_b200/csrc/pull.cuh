// pull.cuh -- DENSE iterations of the level-synchronous push, executed in gather ("pull") form.
//
// One level-synchronous iteration of variant 0 (push.cuh) pops every frontier vertex u -- ru = r[u], r[u] = 0,
// p[u] += a ru -- and adds (1-a) ru / (outdeg(w)+1) to r[w] of every in-neighbour w: one random FP64 atomic per
// traversed edge.  When the frontier covers a large part of the graph the same iteration is cheaper the other
// way round: publish the popped residuals as a dense vector x (x[u] = ru for frontier vertices, 0 elsewhere) and
// let every vertex w GATHER
//     r'[w] = (popped ? 0 : r[w]) + (1-a) / (outdeg(w)+1) * sum_{u in out(w)} x[u]
// from its OUT-list.  No atomics, no frontier queue, no owner search; each (w, source) is written by exactly one
// thread; per edge one 8-byte read of x (a 32-byte sector for 4 sources at once when several sources share the
// launch: x is vertex-major).  The reference has no counterpart (it always scatters, gpu/ExpandRev.cuh); the
// iteration is the same Jacobi step, so the residual bound and the estimates are those of the push form up to the
// order of the floating-point sums.
//
// The pop is deferred by one sweep: sweep k leaves x_next[w] = r'[w] where r'[w] is legal (p[w] is untouched, and
// r[w] is not even written: a non-zero x entry IS the residual while the episode lasts), and sweep k+1 -- or nobody,
// if the loop goes back to scatter mode -- performs p[w] += a x[w], r[w] = 0 for it.  Leaving dense mode is therefore
// just a compaction of the non-zero x entries into an ordinary (un-popped) frontier queue, which also writes them
// back to r.
//
// Work split by out-degree: see pull_sweep.  The longest lists (big_min or more entries) are cut into chunks dealt
// to all CTAs, their partial sums meet in `bigacc` and the vertex is finished after one more grid barrier.
//
// Included by push.cuh (needs PushArgs / PushSmem / the grid barrier).
#pragma once

namespace dppr {

constexpr int kPullBigChunk = 16 * kThreads;     // out-edges per grid-tier chunk

// x is gathered at random and re-read every sweep: keep it in L2 (evict_last), and let everything that merely streams
// through -- out-list slots, ring metadata, r / p rows -- leave first (ld.cs / st.cs).  ncu on the Twitter-shaped graph:
// 50 % L2 hit rate with default policies, DRAM at its random-sector ceiling (~1.3 TB/s).
__device__ __forceinline__ unsigned long long l2_keep_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <class T> __device__ __forceinline__ T pl_ldcs(const T *p);
#ifndef DPPR_PULL_HINTS
#define DPPR_PULL_HINTS 3   // bit 0: streaming (evict-first) loads / stores, bit 1: evict_last gathers of x
#endif
__device__ __forceinline__ double ld_keep(const double *p) {
    if (!(DPPR_PULL_HINTS & 2)) return *p;
    double v;
    asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(l2_keep_policy()));
    return v;
}
__device__ __forceinline__ double2 ld_keep2(const double2 *p) {
    if (!(DPPR_PULL_HINTS & 2)) return *p;
    double2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(l2_keep_policy()));
    return v;
}

template <class T> __device__ __forceinline__ T pl_ldcs(const T *p) { return (DPPR_PULL_HINTS & 1) ? __ldcs(p) : __ldcg(p); }
template <class T> __device__ __forceinline__ void pl_stcs(T *p, T v) { if (DPPR_PULL_HINTS & 1) __stcs(p, v); else __stcg(p, v); }

template <int SB>
__device__ __forceinline__ void pull_gather(const double *x, uint32_t u, uint32_t Sp, uint32_t s0, double (&acc)[SB],
                                            uint32_t &nz) {
    if (SB == 1) {
        const double v = ld_keep(&x[u]);
        acc[0] += v;
        nz += (v != 0.0) ? 1u : 0u;
    } else {
        const double2 *px = reinterpret_cast<const double2 *>(x + (size_t)u * Sp + s0);
#pragma unroll
        for (int j = 0; j < SB / 2; ++j) {
            const double2 v = ld_keep2(&px[j]);
            acc[2 * j] += v.x;
            acc[2 * j + 1] += v.y;
            nz += ((v.x != 0.0) ? 1u : 0u) + ((v.y != 0.0) ? 1u : 0u);
        }
    }
}

template <int SB>
__device__ __forceinline__ void pull_load_x(const double *x, uint32_t w, uint32_t Sp, uint32_t s0, double (&xc)[SB]) {
    if (SB == 1) {
        xc[0] = __ldcg(&x[w]);
    } else {
        const double2 *px = reinterpret_cast<const double2 *>(x + (size_t)w * Sp + s0);
#pragma unroll
        for (int j = 0; j < SB / 2; ++j) {
            const double2 v = __ldcg(&px[j]);
            xc[2 * j] = v.x;
            xc[2 * j + 1] = v.y;
        }
    }
}

// everything that happens once per (vertex, source): deferred pop, new residual, membership in the next frontier
template <int SB>
__device__ __forceinline__ uint32_t pull_finish_unit(const PushArgs &a, int phase, uint32_t w, uint32_t s0, uint32_t len,
                                                     const double (&xc)[SB], const double (&acc)[SB], double *xn,
                                                     unsigned long long &next_edges) {
    uint32_t legal = 0;
    double out[SB];
    const double scale = (1.0 - a.alpha) / (double)(len + 1u);
#pragma unroll
    for (int j = 0; j < SB; ++j) {
        out[j] = 0.0;
        const uint32_t s = s0 + j;
        if (s < (uint32_t)a.S && (len != 0u || xc[j] != 0.0)) {
            // During a dense episode a non-zero x entry IS the residual (r[w] is stale until pull_compact restores it):
            // a vertex that stays in the frontier sweep after sweep -- the steady state -- touches r not at all.
            const size_t idx = (size_t)s * a.Vp + w;
            double rw;
            if (xc[j] != 0.0) {  // w is in the frontier of this sweep: its pop
                pl_stcs(&a.p[idx], pl_ldcs(&a.p[idx]) + a.alpha * xc[j]);
                rw = 0.0;
            } else {
                rw = pl_ldcs(&a.r[idx]);
            }
            rw += acc[j] * scale;
            if (legal_push(rw, phase, a.eps)) {
                out[j] = rw;
                ++legal;
            } else {
                pl_stcs(&a.r[idx], rw);
            }
        }
    }
    if (SB == 1) {
        __stcg(&xn[w], out[0]);
    } else {
        double2 *px = reinterpret_cast<double2 *>(xn + (size_t)w * a.Sp + s0);
#pragma unroll
        for (int j = 0; j < SB / 2; ++j) __stcg(&px[j], make_double2(out[2 * j], out[2 * j + 1]));
    }
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

// Work unit of the dense passes: (vertex w, chunk of SB sources).  With several sources, G = 2^pull_gshift ADJACENT
// lanes take the G chunks of one chunk group of the same vertex: their 32-byte gathers of a neighbour's row then form one
// contiguous G*32-byte request (HBM bursts are 64 bytes: a lone 32-byte sector wastes half of one), and the slot loads
// are a broadcast.  A tile is kThreads / G consecutive vertices x one chunk group.
struct PullUnit {
    uint32_t w, s0, g;
};
template <int SB>
__device__ __forceinline__ uint32_t pull_gshift(const PushArgs &a) { return SB == 1 ? 0u : (uint32_t)a.pull_gshift; }
template <int SB>
__device__ __forceinline__ uint32_t pull_tiles_per_group(const PushArgs &a) {
    const uint32_t vpt = (uint32_t)kThreads >> pull_gshift<SB>(a);
    return ((uint32_t)a.V + vpt - 1) / vpt;
}
template <int SB>
__device__ __forceinline__ PullUnit pull_unit(const PushArgs &a, uint32_t tile, uint32_t tpc, uint32_t &cg) {
    const uint32_t gs = pull_gshift<SB>(a), G = 1u << gs;
    cg = tile / tpc;
    const uint32_t g = threadIdx.x & (G - 1u);
    return PullUnit{(tile - cg * tpc) * ((uint32_t)kThreads >> gs) + (threadIdx.x >> gs), ((cg << gs) + g) * SB, g};
}

// entering dense mode: x[0][w] = r[w] where legal, 0 elsewhere; x[1][w] = 0.  Also lists the ACTIVE tiles: a vertex
// without out-edges receives no adds, so a tile whose 256 vertices have neither out-edges nor a legal residual now
// stays all-zero in both x buffers for the whole episode and is never visited again.  With the degree-sorted internal
// order most vertices of a power-law window sit in such tiles (72 % on the Twitter-shaped window).
// Tiles are visited -- and therefore listed -- in a scrambled order: the heavy tiles (heads of the relabel blocks)
// sit at a regular stride, which would otherwise hand all of them to the same few CTAs in the sweeps.
template <int SB>
__device__ void pull_build(const PushArgs &a, PushSmem &sm, PushCtrl *c, int phase, unsigned int *cnt_out) {
    const uint32_t V = (uint32_t)a.V, nCG = ((uint32_t)a.Sp / SB) >> pull_gshift<SB>(a);
    const uint32_t tpc = pull_tiles_per_group<SB>(a), ntiles = tpc * nCG;
    uint32_t legal = 0;
    unsigned long long ep_slots = 0, ep_pairs = 0, ep_units = 0;  // what ONE sweep over the active tiles moves
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint32_t tile = (uint32_t)(((unsigned long long)t * a.pull_tile_mul) % ntiles);
        uint32_t cg;
        const PullUnit un = pull_unit<SB>(a, tile, tpc, cg);
        const uint32_t w = un.w, s0 = un.s0;
        bool active = false;
        uint32_t wlen = 0;
        if (w < V) {
            double out[SB];
            wlen = (uint32_t)__ldg(&a.outdeg[w]);
            active = wlen != 0;
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                out[j] = 0.0;
                if (s0 + j < (uint32_t)a.S) {
                    const double rw = __ldcg(&a.r[(size_t)(s0 + j) * a.Vp + w]);
                    if (legal_push(rw, phase, a.eps)) { out[j] = rw; ++legal; active = true; }
                }
            }
            if (SB == 1) {
                __stcg(&a.x[0][w], out[0]);
                __stcg(&a.x[1][w], 0.0);
            } else {
                double2 *p0 = reinterpret_cast<double2 *>(a.x[0] + (size_t)w * a.Sp + s0);
                double2 *p1 = reinterpret_cast<double2 *>(a.x[1] + (size_t)w * a.Sp + s0);
#pragma unroll
                for (int j = 0; j < SB / 2; ++j) {
                    __stcg(&p0[j], make_double2(out[2 * j], out[2 * j + 1]));
                    __stcg(&p1[j], make_double2(0.0, 0.0));
                }
            }
        }
        if (__syncthreads_or(active)) {
            if (threadIdx.x == 0) a.tile_list[atomicAdd(&c->ntiles_active, 1u)] = tile;
            if (w < V && s0 < (uint32_t)a.S) {
                const uint32_t nreal = min((uint32_t)SB, (uint32_t)a.S - s0);
                ep_units += nreal;
                ep_pairs += (unsigned long long)wlen * nreal;
                if (un.g == 0) ep_slots += wlen;
            }
        }
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

// one sweep.  `gath` counts the gathered x entries that were non-zero: exactly the (edge, source) pairs the push
// form would have traversed.
//
// Work split by out-degree, per tile of 256 consecutive vertices (consecutive internal ids have similar degrees,
// window.cuh "internal vertex order"):  < warp_min: the owning thread walks its list;  < cta_min: the list goes on a
// shared list whose entries are dealt to the CTA's warps;  < big_min: the whole CTA walks it;  beyond: grid tier.
// (An edge-balanced walk with a segmented warp reduction -- the scatter kernel's scheme -- was measured at 207 us
// per sweep on the Orkut/4 probe against 97 us for this one: the reduction costs more than the imbalance.)
template <int SB>
__device__ void pull_sweep(const PushArgs &a, PushSmem &sm, PushCtrl *c, int phase, const double *xcur, double *xnext,
                           unsigned int *cnt_out, unsigned long long *edges_out, unsigned long long &gath) {
    const uint32_t V = (uint32_t)a.V, Sp = (uint32_t)a.Sp;
    const uint32_t gs = pull_gshift<SB>(a), G = 1u << gs;
    const uint32_t tpc = pull_tiles_per_group<SB>(a), ntiles = __ldcg(&c->ntiles_active);
    uint32_t legal = 0, nz = 0;
    unsigned long long next_edges = 0;
    double *list_res = reinterpret_cast<double *>(sm.stage);   // [kThreads / G][G][SB] sums of the listed vertices (the stage is idle during a sweep)
    double *cta_part = sm.t_ru;                                // [kWarps][G][SB]
    static_assert(kStage >= kThreads * SB && kTileMax >= kWarps * 8 * SB && kTileMax >= kThreads, "pull.cuh borrows the tile arrays");
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint32_t tile = __ldcg(&a.tile_list[t]);  // active tiles only, in the scrambled order pull_build listed them
        uint32_t cg;
        const PullUnit un = pull_unit<SB>(a, tile, tpc, cg);
        const uint32_t w = un.w, s0 = un.s0, g = un.g;
        const bool have = w < V;
        double xc[SB], acc[SB];
#pragma unroll
        for (int j = 0; j < SB; ++j) { xc[j] = 0.0; acc[j] = 0.0; }
        uint32_t len = 0, base = 0, head = 0, mask = 0;
        if (have) {
            len = (uint32_t)pl_ldcs(&a.outdeg[w]);
            pull_load_x<SB>(xcur, w, Sp, s0, xc);
            if (len) {
                const uint4 m = pl_ldcs(&a.vmeta_out[w]);
                base = m.x; head = m.y; len = m.z; mask = m.w - 1u;
            }
        }
        const int tier = len < (uint32_t)a.pull_warp_min ? 0 : len < (uint32_t)a.pull_cta_min ? 1 : len < (uint32_t)a.pull_big_min ? 2 : 3;
        // ---- warp and CTA tiers (one list entry per vertex: the G lanes of a group share it) ----
        const bool listed = tier == 1 || tier == 2;
        if (__syncthreads_or(listed)) {
            uint32_t myslot = 0;
            if (threadIdx.x == 0) sm.pl_n = 0;
            __syncthreads();
            if (listed && g == 0) {
                myslot = atomicAdd(&sm.pl_n, 1u);
                sm.t_base[myslot] = base; sm.t_head[myslot] = head; sm.t_mask[myslot] = mask; sm.t_off[myslot] = len;
            }
            myslot = __shfl_sync(kFull, myslot, lane_id() & ~(G - 1u));
            __syncthreads();
            const uint32_t ne = sm.pl_n;
            for (uint32_t e = warp_id(); e < ne; e += kWarps) {
                const uint32_t el = sm.t_off[e];
                if (el >= (uint32_t)a.pull_cta_min) continue;
                const uint32_t eb = sm.t_base[e], eh = sm.t_head[e], em = sm.t_mask[e];
                double part[SB];
#pragma unroll
                for (int j = 0; j < SB; ++j) part[j] = 0.0;
#pragma unroll 4
                for (uint32_t k = lane_id() >> gs; k < el; k += 32u >> gs)
                    pull_gather<SB>(xcur, (uint32_t)pl_ldcs(&a.pool[eb + ((eh + k) & em)]), Sp, s0, part, nz);
#pragma unroll
                for (int j = 0; j < SB; ++j) {
                    for (uint32_t off = 16; off >= G; off >>= 1) part[j] += __shfl_xor_sync(kFull, part[j], off);
                    if (lane_id() < G) list_res[((e << gs) + g) * SB + j] = part[j];
                }
            }
            for (uint32_t e = 0; e < ne; ++e) {
                const uint32_t el = sm.t_off[e];
                if (el < (uint32_t)a.pull_cta_min) continue;  // (uniform over the CTA)
                const uint32_t eb = sm.t_base[e], eh = sm.t_head[e], em = sm.t_mask[e];
                double part[SB];
#pragma unroll
                for (int j = 0; j < SB; ++j) part[j] = 0.0;
#pragma unroll 4
                for (uint32_t k = threadIdx.x >> gs; k < el; k += (uint32_t)kThreads >> gs)
                    pull_gather<SB>(xcur, (uint32_t)pl_ldcs(&a.pool[eb + ((eh + k) & em)]), Sp, s0, part, nz);
#pragma unroll
                for (int j = 0; j < SB; ++j) {
                    for (uint32_t off = 16; off >= G; off >>= 1) part[j] += __shfl_xor_sync(kFull, part[j], off);
                    if (lane_id() < G) cta_part[((warp_id() << gs) + g) * SB + j] = part[j];
                }
                __syncthreads();
                if (threadIdx.x < G * SB) {  // thread = (sub-chunk, source within it)
                    const uint32_t gg = threadIdx.x / SB, jj = threadIdx.x % SB;
                    double tsum = 0.0;
#pragma unroll
                    for (int ww = 0; ww < kWarps; ++ww) tsum += cta_part[((ww << gs) + gg) * SB + jj];
                    list_res[((e << gs) + gg) * SB + jj] = tsum;
                }
                __syncthreads();
            }
            __syncthreads();
            if (listed) {
#pragma unroll
                for (int j = 0; j < SB; ++j) acc[j] = list_res[((myslot << gs) + g) * SB + j];
            }
            __syncthreads();  // the list is reused by the next tile
        }
        // ---- thread tier ----
        if (tier == 0) {
#pragma unroll 4
            for (uint32_t k = 0; k < len; ++k)
                pull_gather<SB>(xcur, (uint32_t)pl_ldcs(&a.pool[base + ((head + k) & mask)]), Sp, s0, acc, nz);
        }
        // ---- grid tier: finished by pull_big_finish after the next grid barrier ----
        if (tier == 3) {
            if (g == 0) {  // one entry per (vertex, chunk group)
                const uint32_t nch = (len + kPullBigChunk - 1) / kPullBigChunk;
                const unsigned long long old = atomicAdd(&c->bigpk, (1ull << 32) | nch);
                const uint32_t hp = (uint32_t)(old >> 32);
                if (hp < a.bigcap) {
                    __stcg(&a.big[hp].item, ((unsigned long long)cg << 32) | w);
                    __stcg(&a.big[hp].chunk0, (uint32_t)old);
                } else {
                    atomicOr(&c->errflags, kErrHubQ);
                }
            }
        } else if (have) {
            legal += pull_finish_unit<SB>(a, phase, w, s0, len, xc, acc, xnext, next_edges);
        }
    }
    gath += nz;
    pull_count_flush(sm, legal, cnt_out, next_edges, edges_out);
}

// chunks of the grid-tier lists, dealt round-robin to the CTAs
template <int SB>
__device__ void pull_big_expand(const PushArgs &a, PushSmem &sm, const double *xcur, unsigned long long bp,
                                unsigned long long &gath) {
    const uint32_t nh = min((uint32_t)(bp >> 32), a.bigcap), nchunks = (uint32_t)bp;
    const uint32_t Sp = (uint32_t)a.Sp;
    const uint32_t gs = pull_gshift<SB>(a), G = 1u << gs, g = threadIdx.x & (G - 1u);
    double *cta_part = sm.t_ru;
    uint32_t nz = 0;
    for (uint32_t cidx = blockIdx.x; cidx < nchunks; cidx += gridDim.x) {
        uint32_t lo = 0, hi = nh;  // last list entry with chunk0 <= cidx
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldcg(&a.big[mid].chunk0) <= cidx) lo = mid; else hi = mid;
        }
        const unsigned long long item = __ldcg(&a.big[lo].item);
        const uint32_t c0 = __ldcg(&a.big[lo].chunk0);
        const uint32_t w = (uint32_t)item, s0 = ((((uint32_t)(item >> 32)) << gs) + g) * SB;
        const uint4 m = __ldg(&a.vmeta_out[w]);
        const uint32_t e0 = (cidx - c0) * (uint32_t)kPullBigChunk;
        const uint32_t e1 = min(m.z, e0 + (uint32_t)kPullBigChunk);
        double part[SB];
#pragma unroll
        for (int j = 0; j < SB; ++j) part[j] = 0.0;
#pragma unroll 4
        for (uint32_t k = e0 + (threadIdx.x >> gs); k < e1; k += (uint32_t)kThreads >> gs)
            pull_gather<SB>(xcur, (uint32_t)pl_ldcs(&a.pool[m.x + ((m.y + k) & (m.w - 1u))]), Sp, s0, part, nz);
#pragma unroll
        for (int j = 0; j < SB; ++j) {
            for (uint32_t off = 16; off >= G; off >>= 1) part[j] += __shfl_xor_sync(kFull, part[j], off);
            if (lane_id() < G) cta_part[((warp_id() << gs) + g) * SB + j] = part[j];
        }
        __syncthreads();
        if (threadIdx.x < G * SB) {
            const uint32_t gg = threadIdx.x / SB, jj = threadIdx.x % SB;
            double t = 0.0;
#pragma unroll
            for (int ww = 0; ww < kWarps; ++ww) t += cta_part[((ww << gs) + gg) * SB + jj];
            atomicAdd(&a.bigacc[(((size_t)lo << gs) + gg) * 4 + jj], t);
        }
        __syncthreads();
    }
    gath += nz;
}

template <int SB>
__device__ void pull_big_finish(const PushArgs &a, PushSmem &sm, int phase, const double *xcur, double *xnext,
                                unsigned long long bp, unsigned int *cnt_out, unsigned long long *edges_out) {
    const uint32_t nh = min((uint32_t)(bp >> 32), a.bigcap);
    const uint32_t gs = pull_gshift<SB>(a), G = 1u << gs;
    uint32_t legal = 0;
    unsigned long long next_edges = 0;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < (nh << gs); i += gridDim.x * kThreads) {
        const uint32_t h = i >> gs, g = i & (G - 1u);  // thread = (list entry, sub-chunk)
        const unsigned long long item = __ldcg(&a.big[h].item);
        const uint32_t w = (uint32_t)item, s0 = ((((uint32_t)(item >> 32)) << gs) + g) * SB;
        double xc[SB], acc[SB];
        pull_load_x<SB>(xcur, w, (uint32_t)a.Sp, s0, xc);
#pragma unroll
        for (int j = 0; j < SB; ++j) {
            acc[j] = __ldcg(&a.bigacc[(size_t)i * 4 + j]);
            __stcg(&a.bigacc[(size_t)i * 4 + j], 0.0);
        }
        const uint32_t len = __ldg(&a.vmeta_out[w]).z;
        legal += pull_finish_unit<SB>(a, phase, w, s0, len, xc, acc, xnext, next_edges);
    }
    pull_count_flush(sm, legal, cnt_out, next_edges, edges_out);
}

// leaving dense mode: the non-zero entries of x are the (un-popped) frontier of the next scatter iteration
template <int SB>
__device__ void pull_compact(const PushArgs &a, PushSmem &sm, PushCtrl *c, const double *x, unsigned long long *qout,
                             unsigned int *cnt_out) {
    const uint32_t V = (uint32_t)a.V;
    const uint32_t tpc = pull_tiles_per_group<SB>(a), ntiles = __ldcg(&c->ntiles_active);
    for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint32_t tile = __ldcg(&a.tile_list[t]);
        uint32_t cg;
        const PullUnit un = pull_unit<SB>(a, tile, tpc, cg);
        const uint32_t w = un.w, s0 = un.s0;
        double xc[SB];
#pragma unroll
        for (int j = 0; j < SB; ++j) xc[j] = 0.0;
        if (w < V) pull_load_x<SB>(x, w, (uint32_t)a.Sp, s0, xc);
#pragma unroll
        for (int j = 0; j < SB; ++j) {
            if (xc[j] != 0.0) __stcg(&a.r[(size_t)(s0 + j) * a.Vp + w], xc[j]);  // (see pull_finish_unit)
            stage_push(xc[j] != 0.0, ((unsigned long long)(s0 + j) << 32) | w, sm, qout, cnt_out, a.qcap, a.ctrl);
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
        c->ep_slots = 0; c->ep_pairs = 0; c->ep_units = 0;
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
            const size_t idx = (size_t)(item >> 32) * a.Vp + (uint32_t)item;
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
    while (true) {
        const uint32_t n = __ldcg(&c->dcnt[k % 3]);
        if (n == 0) break;
        if (k > 0) {
            const unsigned long long ne = __ldcg(&c->dedges[k % 3]);
            const bool leave = (rate > 0.f && sw_known > 0.f) ? (double)ne * (double)rate < 0.8 * (double)sw_known
                                                               : ne < a.dense_exit_edges;
            if (leave) break;
        }
        if ((int)iters_done >= a.max_iters) {
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
        pull_sweep<SB>(a, sm, c, phase, a.x[cur], a.x[cur ^ 1], &c->dcnt[(k + 1) % 3], &c->dedges[(k + 1) % 3], gath);
        if (!grid_barrier(c, gen, sm)) return false;
        const unsigned long long bp = __ldcg(&c->bigpk);
        if (bp) {
            pull_big_expand<SB>(a, sm, a.x[cur], bp, gath);
            if (!grid_barrier(c, gen, sm)) return false;
            if (blockIdx.x == 0 && threadIdx.x == 0) c->bigpk = 0;
            pull_big_finish<SB>(a, sm, phase, a.x[cur], a.x[cur ^ 1], bp, &c->dcnt[(k + 1) % 3], &c->dedges[(k + 1) % 3]);
            if (!grid_barrier(c, gen, sm)) return false;
        }
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
    if (__ldcg(&c->dcnt[k % 3]) != 0)
        pull_compact<SB>(a, sm, c, a.x[cur], a.q[(it + 1) & 1], &c->cnt[(it + 1) % 3]);
    return grid_barrier(c, gen, sm);
}

// (not inlined, and its in/out scalars travel in one struct: the scatter path keeps its register allocation)
struct DenseIO {
    unsigned long long edges_acc, gath, pops_acc;
    uint32_t iters_done, sweeps_done;
    GridBar gen;
    float rate;  // scatter cost per edge the caller decided with (0 = not measured yet)
};
template <int SB>
__device__ __noinline__ bool dense_mode(const PushArgs a, PushSmem &sm, PushCtrl *c, int phase, uint32_t it,
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
