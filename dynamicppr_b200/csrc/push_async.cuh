// push_async.cuh -- ticket-queue reverse push for variant 0 (OPTIMIZED = eager + fast frontier).
//
// Measurements of the CTA-tiled level-synchronous kernel in push.cuh (profiles/README.md) show that on an
// L2-resident graph the push is bound by chains of dependent L2 round trips (0.19 us per load, 0.44 us per
// returning FP64 atomic, 2.1 us per 592-CTA grid barrier: profiles/micro_latency_r01.txt), that whole CTAs
// convoy behind their slowest warp at every __syncthreads, and that every iteration waits for the slowest
// tile (median CTA done at 12 us, last one at 31 us).  This engine keeps the level-synchronous SCHEDULE
// (a vertex accumulates all adds of a generation before it is popped -- popping eagerly at the first
// crossing was measured to inflate the work 2.2x) but removes CTA tiles, __syncthreads and grid barriers:
//
//   * ONE device-wide ticket ring per phase.  `tail` = items ever enqueued, `head` = tickets ever handed
//     out, `done` = items completely processed, `fence` = (generation << 48) | first ticket of the NEXT
//     generation.  A warp always holds 32 tickets (one per lane); a lane may consume its ticket once it is
//     below the fence.  Items enqueued while generation g is processed land at or beyond the fence, i.e.
//     in generation g+1.  Work is dealt 32 items at a time to whichever warp is free: dynamic balance.
//   * A warp processes its ready lanes together: claim r[u] with one atomicExch (every later add is kept
//     and re-enqueues u exactly when it crosses the threshold -- the variant-0 rule of push.cuh), RED into
//     p[u], warp-scan of the in-degrees, 4 edges per lane per round with all loads of a stage in flight,
//     threshold crossings staged in per-warp shared memory, ONE atomicAdd on `tail` per warp pass, then
//     `done` is bumped with release semantics.
//   * Generation switch without a barrier: an idle lane 0 that sees done == fence reads `tail` and CASes
//     the fence word to (generation + 1, tail); done == fence == tail ends the phase.
//   * No fences on the hot path.  gpu-scope acquire/release make ptxas emit MEMBAR + CCTL.IVALL (whole-L1
//     invalidate); with thousands of polling warps that was 31 % of all stall samples (profiles/README.md).
//     Instead every cross-warp hand-off is self-validating: a slot holds EMPTY until its item is stored and
//     the consumer spins on it (relaxed, L2) and writes EMPTY back; `done` is bumped after the warp's `tail`
//     reservation RETURNED (data dependence), so "done == fence" implies `tail` is final.
//   * Carry-over thresholds: generation g only pushes items with |r| > theta_g = max(eps, theta_0*gamma^g);
//     smaller ones are put back and carried to the next generation untouched.  Vertices then collect more
//     mass before they are pushed: ~43 % fewer in-edge traversals for ~45 % more (cheap) generations
//     (scripts/experiments/carry_sim.c).  The final generations run with theta = eps, so the result obeys
//     the same contract as every other variant: all |r| <= eps.
//   * Vertices with in-degree >= hub_degree are split into 1024-edge chunk items (processed in the next
//     generation) carrying the claimed residual in a side array.
//   * The two phases use separate rings and counters: one grid barrier per phase (after seeding).
#pragma once
#include "push.cuh"

namespace dppr {

constexpr unsigned long long kEmptySlot = ~0ull;
constexpr int kAsyncStage = 128;                 // staged enqueues per warp
constexpr unsigned long long kChunkFlag = 1ull << 63;
constexpr int kChunkShift = 44;                  // item = flag(1) | chunk(19) | source(12) | vertex(32)
constexpr int kMaxAsyncSources = 4096;

struct AsyncQueue {
    unsigned long long *slots;   // ring, all EMPTY when idle
    double *slot_ru;             // claimed residual of chunk items (same indexing)
    unsigned long long mask;     // capacity - 1
    unsigned long long *tail;    // each counter on its own 128-byte line
    unsigned long long *head;
    unsigned long long *done;
    unsigned long long *fence;   // (generation << 48) | first ticket of the next generation
    unsigned long long *theta0;  // bits of max |r| over the seeds (positive double => integer order)
};

struct AsyncArgs {
    PushArgs base;
    AsyncQueue q[2];             // one ring per phase
    int32_t guard_slots;         // unused (kept for ABI stability of the launch struct)
    double carry_gamma;          // theta_g = max(eps, theta0 * carry_scale * gamma^g); gamma >= 1 disables carrying
    double carry_scale;
    unsigned long long *dbg;     // optional debug counters: passes, items, pass cycles, idle polls, slot spins, rounds
};


struct AsyncWarpSmem {
    unsigned long long stage[kAsyncStage];
    double ru[32];
    uint32_t off[33];
    uint32_t base[32];
    uint32_t head[32];
    uint32_t mask[32];
    uint32_t s[32];
    uint32_t stage_cnt;
};

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- enqueue: the warp's staged items -> ring.  Called by all 32 lanes. -------------------------
__device__ __forceinline__ void async_flush(const AsyncArgs &a, const AsyncQueue &q, AsyncWarpSmem &ws) {
    __syncwarp();
    const uint32_t n = ws.stage_cnt;
    if (n == 0) return;
    unsigned long long pos = 0;
    if (lane_id() == 0) pos = atomicAdd(q.tail, (unsigned long long)n);
    pos = __shfl_sync(kFull, pos, 0);
    for (uint32_t i = lane_id(); i < n; i += 32) st_relaxed_u64(&q.slots[(pos + i) & q.mask], ws.stage[i]);
    __syncwarp();
    if (lane_id() == 0) ws.stage_cnt = 0;
    __syncwarp();
}

// stage one item per wanting lane; flushes first when the stage could overflow.  All 32 lanes call.
__device__ __forceinline__ void async_stage(const AsyncArgs &a, const AsyncQueue &q, AsyncWarpSmem &ws, bool want,
                                            unsigned long long item) {
    const unsigned m = __ballot_sync(kFull, want);
    if (m == 0) return;
    if (ws.stage_cnt + __popc(m) > (unsigned)kAsyncStage) async_flush(a, q, ws);
    if (want) ws.stage[ws.stage_cnt + __popc(m & lanemask_lt())] = item;
    __syncwarp();
    if (lane_id() == 0) ws.stage_cnt += __popc(m);
    __syncwarp();
}

// chunk items are written straight to the ring (they carry a residual that must be visible first)
__device__ __forceinline__ void async_enqueue_chunks(const AsyncArgs &a, const AsyncQueue &q, unsigned long long item,
                                                     double ru, uint32_t nchunks) {
    const unsigned long long pos = atomicAdd(q.tail, (unsigned long long)nchunks);
    for (uint32_t c = 0; c < nchunks; ++c) {
        const unsigned long long slot = (pos + c) & q.mask;
        __stcg(&q.slot_ru[slot], ru);
        st_release_u64(&q.slots[slot], item | kChunkFlag | ((unsigned long long)c << kChunkShift));  // hubs only
    }
}

// ---- edges of one round: 32 lanes x kEdgeUnroll ---------------------------------------------------
template <class Owner>
__device__ __forceinline__ void async_push_edges(const AsyncArgs &a, const AsyncQueue &q, AsyncWarpSmem &ws,
                                                 const Owner &ow, const uint32_t (&nbr)[kEdgeUnroll],
                                                 const bool (&active)[kEdgeUnroll], int phase) {
    const PushArgs &b = a.base;
    int32_t dv[kEdgeUnroll];
    double add[kEdgeUnroll], old[kEdgeUnroll];
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k) dv[k] = active[k] ? __ldg(&b.outdeg[nbr[k]]) : 0;
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k) {
        old[k] = 0.0; add[k] = 0.0;
        if (active[k]) {
            add[k] = ow.ru_scaled(k) / (double)(dv[k] + 1);
            old[k] = atomicAdd(&b.r[ow.sb(k) + nbr[k]], add[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k) {
        const bool want = active[k] && !legal_push(old[k], phase, b.eps) && legal_push(old[k] + add[k], phase, b.eps);
        async_stage(a, q, ws, want, ((unsigned long long)ow.s(k) << 32) | nbr[k]);
    }
}

struct AsyncTileOwner {
    const AsyncWarpSmem &ws;
    uint32_t lo[kEdgeUnroll];
    long long Vp;
    __device__ __forceinline__ double ru_scaled(int k) const { return ws.ru[lo[k]]; }
    __device__ __forceinline__ unsigned long long sb(int k) const { return (unsigned long long)ws.s[lo[k]] * Vp; }
    __device__ __forceinline__ uint32_t s(int k) const { return ws.s[lo[k]]; }
};

// ---- seeds: warp-strided scan of the candidates, straight into the ring ---------------------------------
__device__ void async_seed(const AsyncArgs &a, const AsyncQueue &q, AsyncWarpSmem &ws, int phase) {
    const PushArgs &b = a.base;
    const uint32_t ncand = b.init_mode ? 1u : __ldcg(b.ncand);
    const unsigned long long total = (unsigned long long)ncand * (unsigned)b.S;
    const unsigned long long stride = (unsigned long long)gridDim.x * kThreads;
    const unsigned long long rounds = (total + stride - 1) / stride;
    double mx = 0.0;
    for (unsigned long long rd = 0; rd < rounds; ++rd) {
        const unsigned long long j = rd * stride + (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
        bool want = false;
        unsigned long long item = 0;
        if (j < total) {
            const uint32_t s = (uint32_t)(j / ncand);
            const uint32_t c = (uint32_t)(j - (unsigned long long)s * ncand);
            const uint32_t u = b.init_mode ? (uint32_t)b.src[s] : b.cand[c];
            const double x = __ldcg(&b.r[(unsigned long long)s * b.Vp + u]);
            want = legal_push(x, phase, b.eps);
            item = ((unsigned long long)s << 32) | u;
            if (want) mx = fmax(mx, fabs(x));
        }
        async_stage(a, q, ws, want, item);
    }
    async_flush(a, q, ws);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(kFull, mx, off));
    if (lane_id() == 0 && mx > 0.0) atomicMax(q.theta0, (unsigned long long)__double_as_longlong(mx));
}

// ---- the worker loop of one warp for one phase ------------------------------------------------------------
__device__ __forceinline__ unsigned long long atom_cas_acq_rel_u64(unsigned long long *p, unsigned long long cmp,
                                                                   unsigned long long val) {
    unsigned long long old;
    asm volatile("atom.acq_rel.gpu.global.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "l"(p), "l"(cmp), "l"(val) : "memory");
    return old;
}

// The fence word is written by exactly one warp (the switcher: block 0, warp 0, which processes no items), so
// `done` and `tail` have a single reader and the thousands of RED updates on `done` never queue behind polls.
// Bit 47 of the fence word = phase over.
constexpr unsigned long long kFenceOver = 1ull << 47;
constexpr unsigned long long kFencePos = kFenceOver - 1;

__device__ void async_switcher(const AsyncArgs &a, const AsyncQueue &q, int phase, unsigned long long &gens_out) {
    const PushArgs &b = a.base;
    unsigned long long fw = ld_relaxed_u64(q.tail);          // generation 0 = the seeds (tail is quiescent here)
    unsigned gen = 0;
    if (lane_id() == 0) {
        const long long t0 = clock64();
        long long t_last = t0;
        unsigned long long d_last = ~0ull;
        while (true) {
            if (a.dbg && phase == 0 && gen == 20) a.dbg[8] = global_ns();
            st_relaxed_u64(q.fence, ((unsigned long long)gen << 48) | fw | (1ull << 46));  // bit 46: "published"
            unsigned long long d;
            unsigned spins = 0;
            while ((d = ld_relaxed_u64(q.done)) != fw) {
                if (d != d_last) { d_last = d; t_last = clock64(); }
                else if ((++spins & 0x3ffu) == 0 && clock64() - t_last > 3000000000ll) {
                    // ~1.5 s without a single item completing: give up loudly instead of hanging the GPU
                    atomicOr(&b.ctrl->errflags, kErrWatchdog);
                    printf("[dppr] push watchdog: phase %d gen %u done %llu fence %llu tail %llu head %llu\n", phase, gen, d, fw,
                           ld_relaxed_u64(q.tail), ld_relaxed_u64(q.head));
                    goto out;
                }
            }
            if (a.dbg && phase == 0 && gen == 20) { a.dbg[9] = global_ns(); a.dbg[14] = fw; }
            // every item below the fence is processed, so nobody is producing: tail is final for this generation
            const unsigned long long t = ld_relaxed_u64(q.tail);
            if (t == fw) break;
            fw = t;
            ++gen;
        }
    out:
        st_relaxed_u64(q.fence, ((unsigned long long)gen << 48) | fw | kFenceOver | (1ull << 46));
    }
    gens_out = __shfl_sync(kFull, gen, 0);
}

__device__ void async_worker(const AsyncArgs &a, const AsyncQueue &q, AsyncWarpSmem &ws, unsigned long long *cta_fence,
                             int phase, unsigned long long &edges_acc, unsigned long long &pops_acc,
                             unsigned long long &hubs_acc) {
    const PushArgs &b = a.base;
    const unsigned lane = lane_id();
    unsigned pend = 0;                 // lanes whose ticket has not been consumed yet (warp-uniform)
    unsigned long long ticket = 0;
    unsigned idle = 0;
    unsigned long long dbg_pass = 0, dbg_items = 0, dbg_cyc = 0, dbg_idle = 0, dbg_spin = 0, dbg_rounds = 0;
    const double theta0 = __longlong_as_double((long long)ld_relaxed_u64(q.theta0)) * a.carry_scale;
    const bool carrying = a.carry_gamma < 1.0 && a.carry_gamma > 0.0;
    while (true) {
        if (pend == 0) {
            unsigned long long h = 0;
            if (lane == 0) h = atomicAdd(q.head, 32ull);
            ticket = __shfl_sync(kFull, h, 0) + lane;
            pend = kFull;
        }
        // the CTA shares one cached copy of the fence word; an idle warp refreshes it from global memory
        unsigned long long fw = 0;
        if (lane == 0) {  // fence words only grow (generation and position are monotone), so the cache is a max
            fw = atomicMax(cta_fence, 0ull);
            if (idle != 0 && (idle & 1u)) {
                const unsigned long long g = ld_relaxed_u64(q.fence);
                if (g > fw) { fw = g; atomicMax(cta_fence, g); }
            }
        }
        fw = __shfl_sync(kFull, fw, 0);
        const unsigned long long fpos = fw & ((1ull << 46) - 1);
        const unsigned gen = (unsigned)(fw >> 48);
        const bool ready = (pend & (1u << lane)) && ticket < fpos;
        const unsigned m = __ballot_sync(kFull, ready);
        if (m == 0) {
            if (fw & kFenceOver) {
                if (a.dbg && lane == 0) {
                    atomicAdd(&a.dbg[0], dbg_pass); atomicAdd(&a.dbg[1], dbg_items); atomicAdd(&a.dbg[2], dbg_cyc);
                    atomicAdd(&a.dbg[3], dbg_idle); atomicAdd(&a.dbg[4], dbg_spin); atomicAdd(&a.dbg[5], dbg_rounds);
                }
                return;
            }
            ++idle; ++dbg_idle;
            if (idle > 4) __nanosleep(idle < 64 ? 50 : 150);
            continue;
        }
        idle = 0;
        pend &= ~m;
        const long long dbg_t0 = clock64();
        ++dbg_pass; dbg_items += __popc(m);
        if (a.dbg && phase == 0 && gen == 20 && lane == 0) { const unsigned long long t = global_ns(); atomicMin(&a.dbg[10], t); atomicMax(&a.dbg[11], t); atomicAdd(&a.dbg[15], 1ull); }
        double theta = b.eps;
        if (carrying) theta = fmax(b.eps, theta0 * pow(a.carry_gamma, (double)gen));

        // ---- pop (normal items) / unpack (chunk items) ----
        uint32_t deg = 0, ebase = 0;     // edges [ebase, ebase + deg) of the in-list belong to this lane's item
        bool carry = false;
        unsigned long long item = 0;
        if (ready) {
            unsigned long long *slot = &q.slots[ticket & q.mask];
            while ((item = ld_relaxed_u64(slot)) == kEmptySlot) { ++dbg_spin; }  // store still in flight: a few hundred ns at most
            st_relaxed_u64(slot, kEmptySlot);
            const bool is_chunk = (item & kChunkFlag) != 0;
            const uint32_t v = (uint32_t)item, s = (uint32_t)(item >> 32) & (kMaxAsyncSources - 1);
            const unsigned long long idx = (unsigned long long)s * b.Vp + v;
            const uint4 mt = __ldg(&b.vmeta[v]);
            double ru;
            if (!is_chunk) {
                ru = __longlong_as_double((long long)atomicExch((unsigned long long *)&b.r[idx], 0ull));
                if (fabs(ru) <= theta && fabs(ru) > b.eps) {
                    atomicAdd(&b.r[idx], ru);        // RED: put it back, carry the vertex to the next generation
                    carry = true;
                } else if (ru != 0.0) {  // (an exact zero is a duplicate whose twin already took everything)
                    atomicAdd(&b.p[idx], b.alpha * ru);  // RED
                    deg = mt.z;
                    if (deg >= (uint32_t)b.hub_degree) {
                        async_enqueue_chunks(a, q, item, ru, (deg + kHubChunk - 1) / kHubChunk);
                        deg = 0;
                        ++hubs_acc;
                    }
                    ++pops_acc;
                }
            } else {
                ru = __ldcg(&q.slot_ru[ticket & q.mask]);
                const uint32_t c = (uint32_t)(item >> kChunkShift) & 0x7ffffu;
                ebase = c * (uint32_t)kHubChunk;
                deg = min(mt.z - ebase, (uint32_t)kHubChunk);
            }
            ws.ru[lane] = (1.0 - b.alpha) * ru;
            ws.base[lane] = mt.x;
            ws.head[lane] = mt.y + ebase;
            ws.mask[lane] = mt.w - 1u;
            ws.s[lane] = s;
        }
        async_stage(a, q, ws, carry, item);
        // ---- exclusive prefix of the degrees over the warp ----
        const uint32_t inc = warp_inclusive_sum(deg);
        ws.off[lane] = inc - deg;
        const uint32_t total = __shfl_sync(kFull, inc, 31);
        __syncwarp();
        // ---- edges: 32 x kEdgeUnroll per round ----
        for (uint32_t e0 = 0; e0 < total; e0 += 32 * kEdgeUnroll) {
            AsyncTileOwner ow{ws, {}, b.Vp};
            uint32_t nbr[kEdgeUnroll];
            bool active[kEdgeUnroll];
#pragma unroll
            for (int k = 0; k < kEdgeUnroll; ++k) {
                const uint32_t e = e0 + k * 32 + lane;
                active[k] = e < total;
                nbr[k] = 0;
                if (active[k]) {
                    uint32_t lo = 0, hi = 32;  // last j with off[j] <= e
#pragma unroll
                    for (int step = 0; step < 5; ++step) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (ws.off[mid] <= e) lo = mid; else hi = mid;
                    }
                    ow.lo[k] = lo;
                    nbr[k] = (uint32_t)__ldg(&b.pool[ws.base[lo] + ((ws.head[lo] + (e - ws.off[lo])) & ws.mask[lo])]);
                }
            }
            async_push_edges(a, q, ws, ow, nbr, active, phase);
            ++dbg_rounds;
        }
        if (lane == 0) edges_acc += total;
        // ---- publish: enqueues first, then `done` (release) ----
        async_flush(a, q, ws);
        if (lane == 0) red_relaxed_add_u64(q.done, (unsigned long long)__popc(m));  // after the tail reservation returned
        __syncwarp();
        dbg_cyc += clock64() - dbg_t0;
        if (a.dbg && phase == 0 && gen == 20 && lane == 0) { const unsigned long long t = global_ns(); atomicMin(&a.dbg[12], t); atomicMax(&a.dbg[13], t); }
    }
}

__global__ void __launch_bounds__(kThreads, DPPR_MIN_BLOCKS) push_async(const AsyncArgs a) {
    __shared__ AsyncWarpSmem wsm[kWarps];
    struct BarSmem { int abort_flag; uint32_t bar_units; };
    __shared__ BarSmem bsm;
    __shared__ unsigned long long cta_fence;
    AsyncWarpSmem &ws = wsm[warp_id()];
    if (lane_id() == 0) ws.stage_cnt = 0;
    if (threadIdx.x == 0) bsm.abort_flag = 0;
    __syncthreads();
    PushCtrl *c = a.base.ctrl;
    GridBar gen;
    unsigned long long edges_acc = 0, pops_acc = 0, hubs_acc = 0, gens_acc = 0;
    const int nphases = a.base.init_mode ? 1 : 2;
    for (int phase = 0; phase < nphases; ++phase) {
        async_seed(a, a.q[phase], ws, phase);
        if (!grid_barrier(c, gen, bsm)) break;   // every seed is in the ring and theta0 is final
        if (threadIdx.x == 0) cta_fence = 0;
        __syncthreads();
        if (blockIdx.x == 0 && warp_id() == 0) {
            unsigned long long g = 0;
            async_switcher(a, a.q[phase], phase, g);
            gens_acc += g + 1;
        } else {
            async_worker(a, a.q[phase], ws, &cta_fence, phase, edges_acc, pops_acc, hubs_acc);
        }
        __syncthreads();
    }
    // per-lane counters -> warp -> control block
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        pops_acc += __shfl_xor_sync(kFull, pops_acc, off);
        hubs_acc += __shfl_xor_sync(kFull, hubs_acc, off);
    }
    if (lane_id() == 0) {
        if (edges_acc) atomicAdd(&c->edges, edges_acc);
        if (pops_acc) atomicAdd(&c->pops, pops_acc);
        if (hubs_acc) atomicAdd(&c->hubs, hubs_acc);
        if (blockIdx.x == 0 && warp_id() == 0) c->iters = gens_acc;
    }
}

}  // namespace dppr
