// push.cuh -- the reverse-push loop (north-star subsystem 3).
//
// Replaces InspectPureRev / InspectExtra (gpu/Inspect.cuh), UpdateFrontierStatus
// (gpu/PPRCommon.cuh:24-34), the four Expand*Rev kernels and RepairFrontierRev (gpu/ExpandRev.cuh)
// and the host while(1) around them (gpu/PPRRevPushGPU.cuh:97-131, PPRRevPushGPUVariants.cuh),
// which costs a blocking 4-byte D2H + a memset + 2-3 launches per push iteration.
//
// Structure here:
//   * The whole refresh (seed phase 0, iterate to exhaustion, seed phase 1, iterate) is ONE
//     persistent cooperative launch; iterations are separated by a software grid barrier, frontier
//     counters live on the device (three rotating slots so that a slot is only cleared one full
//     iteration after its last reader).  A stepwise mode drives the same device functions with one
//     launch per sub-pass, for debugging and per-iteration profiling.
//   * Seeds come from the repaired vertices only -- the reference GPU path scans all V
//     residuals twice per batch (gpu/Inspect.cuh:9-48).
//   * Frontier items are (source, vertex) pairs, so S sources share one launch, one queue and one
//     window graph.
//   * Neighbour expansion is edge-balanced: a CTA prefix-sums the in-degrees of a tile of frontier
//     items in shared memory and its threads walk the concatenated edge range, finding the owner by
//     binary search, so a thread's work does not depend on its vertex's degree (the reference's
//     CTA / warp / scan three-tier scheme, gpu/ExpandRev.cuh:48-179).  Vertices with in-degree >=
//     hub_degree are moved to a hub list and expanded by the whole grid in the next iteration.
//   * The next frontier is compacted with warp ballot + popc into a shared-memory staging buffer and
//     flushed with one global atomicAdd per CTA per tile (reference: a cub::BlockScan and a global
//     atomic per 256-edge step).
//   * Residual scatter uses the native FP64 atomicAdd returning the old value (the reference emulates
//     it with a CAS loop, gpu/GPUUtil.cuh:21-30): every variant needs `old` for its enqueue rule.
//   * On large windows variant 0 runs the iterations whose frontier covers most of the graph as gather
//     sweeps over the out-lists instead (pull.cuh, instantiation push_persistent<0, true>); the switch is
//     decided on the device from measured costs.
//
// Variants (-o), semantics of SURVEY A.5:
//   0 OPTIMIZED      eager + fast frontier.  ru is claimed with one atomicExch(r[u], 0) at pop time, so
//                    adds that reach u later in the same iteration are kept and u re-enters the next
//                    frontier exactly when they carry it across the threshold -- the same set the
//                    reference obtains by reading r[u] live, subtracting ru afterwards and re-checking
//                    (RepairFrontierRev, gpu/ExpandRev.cuh:709-743), without the second pass.
//   1 FAST_FRONTIER  snapshot pass (InspectExtra: ft_r=r, p+=a r, r=0), barrier, push; threshold-crossing dedupe.
//   2 EAGER          stamp status[u]=level, barrier, live read + push with atomicExch(status) dedupe,
//                    barrier, repair pass r[u]-=ru with re-enqueue.
//   3 VANILLA        snapshot pass, barrier, push with atomicExch(status) dedupe.
#pragma once
#include "common.cuh"
#include "window.cuh"

namespace dppr {

#ifndef DPPR_STAGE
#define DPPR_STAGE 1024
#endif
constexpr int kStage = DPPR_STAGE;      // staged next-frontier items per CTA
#ifndef DPPR_MIN_BLOCKS
#define DPPR_MIN_BLOCKS 4
#endif
#ifndef DPPR_EDGE_UNROLL
#define DPPR_EDGE_UNROLL 4
#endif
constexpr int kEdgeUnroll = DPPR_EDGE_UNROLL;          // in-edges per thread per round: independent load/atomic chains in flight
constexpr int kHubChunk = kEdgeUnroll * kThreads;  // edges of a hub one CTA takes at a time
#ifndef DPPR_HUB_SMEM
#define DPPR_HUB_SMEM 1024
#endif
constexpr int kHubSmem = DPPR_HUB_SMEM;  // hub chunk offsets cached in shared memory for the owner search

// device-resident control block.  [0, kCtrlZeroBytes) is cleared before every refresh.
struct PushCtrl {
    unsigned int cnt[3];    // frontier sizes, slot it % 3 is consumed in iteration `it`
    unsigned int pad1;
    unsigned long long hpk[3];  // hub lists, slot it % 3 is produced in iteration `it`: (hubs << 32) | edge chunks
    unsigned long long bar64[2];  // grid barrier words, used alternately: low half = arrivals (monotone within a launch),
                                  // high half = running sum of the work units the arriving CTAs report
    unsigned long long iters, pops, edges, hubs;
    unsigned long long theta0[2];  // bits of max |r| over the seeds of phase 0 / 1 (positive doubles order like integers)
    unsigned long long carried;    // frontier items carried over untouched (variant 0 threshold schedule)
    unsigned long long dedges[3];  // dense mode (pull.cuh): in-edges a scatter iteration over the next frontier would traverse
    unsigned int dcnt[3];          // dense mode: frontier sizes, rotating like cnt
    unsigned int sweeps;           // dense sweeps of this refresh
    unsigned int ntiles_active;    // dense mode: active tiles of the running episode = sum of ntiles_b
    unsigned int ntiles_b[3];      // ... by weight class (heavy first): tile_list holds three lists of tile_list_cap entries
    unsigned int work_next[2];     // dense mode: next work item (grid-tier chunk, then tile) of the running sweep, by sweep parity
    unsigned int nalt[2];          // dense mode: items in qalt -- residuals of the OTHER sign an episode left behind, the seeds of the
                                   // next phase.  Indexed by the parity of the phase that CONSUMES them: the word a phase reads is never
                                   // the one being reset or appended to between the same two grid barriers
    unsigned int pad0;
    unsigned long long bigpk;      // dense mode: grid-tier list of the running sweep, (entries << 32) | chunks
    unsigned long long gath;       // dense sweeps: gathered x entries that were non-zero = the (edge, source) pairs a scatter
                                   // iteration would have traversed
    unsigned long long walk_slots, walk_pairs, units;  // dense sweeps, what they actually moved: out-list entries walked (per
                                   // chunk group), (entry, source) gathers, (vertex, source) units finished
    unsigned long long ep_slots, ep_pairs, ep_units;   // the same for ONE sweep of the running dense episode (pull_build)
    unsigned long long dpops;      // frontier pops performed by sweeps (included in `pops`)
    // ---- persistent across launches ----
    int errflags;
    int level;              // last status stamp handed out (variants 2, 3)
    // self-calibration of the scatter / gather switch (DENSE kernel): both measured on the device with %globaltimer
    float sweep_ns;         // duration of one gather sweep in the last dense episode (0 = none measured yet)
    float rate_ns[2];       // scatter cost per traversed in-edge, written in slot it & 1 during iteration `it`
    float pad2;
};
constexpr size_t kCtrlZeroBytes = offsetof(PushCtrl, errflags);

struct HubItem {            // 32 bytes
    unsigned long long item;
    double ru;
    uint32_t chunk0;        // index of this hub's first edge chunk in the iteration-wide chunk numbering
    uint32_t pad[3];
};

struct PushArgs {
    const uint4 *vmeta;
    const int32_t *pool;
    const int32_t *outdeg;
    double *p;
    double *r;
    float *pacc;                 // (kernels that can switch, several sources) per pair: sum of the amounts the running sweep episode
                                 // has popped so far; folded into p when the episode ends, zero otherwise (pull.cuh, pull_fold)
    int32_t *status;
    int64_t Sr;                  // row stride of p / r / status / x: these arrays are VERTEX-major, [V][Sr], Sr = 1 or S rounded up
                                 // to a multiple of 8 (padding columns stay zero): element (source s, vertex v) sits at v * Sr + s
    int32_t S;
    const int32_t *src;
    unsigned long long *q[2];
    unsigned long long *qalt;    // (dense kernels) frontier items of the other sign an episode left behind, see pull_compact
    double *qr[2];
    uint32_t qcap;
    HubItem *hub[2];
    uint32_t hcap;
    const uint32_t *cand;        // seed candidates: distinct repaired vertices
    const uint32_t *ncand;       // device scalar
    PushCtrl *ctrl;
    double eps, alpha;
    int32_t hub_degree;
    int32_t init_mode;           // 1: seed = the sources themselves, phase 0 only (initial solve)
    int32_t max_iters;
    uint4 *iterlog;              // debug: per iteration (frontier size, hubs, globaltimer lo, hi) of the last refresh
    int32_t iterlog_cap;
    unsigned long long *ctalog;  // debug: [grid][8] globaltimer stamps of iteration `probe_iter`
    int32_t probe_iter;
    double carry_gamma;          // variant 0: iteration k of a phase only pushes items with |r| > max(eps, theta0*scale*gamma^k);
    double carry_scale;          // the rest is put back and carried to the next frontier.  gamma >= 1 disables carrying.
    int32_t tile_cap;            // tile size (frontier items) once a CTA's share of the frontier exceeds 512 items
    // ---- dense iterations in gather form (pull.cuh), variant 0 ----
    int32_t V;
    float avg_indeg;             // E_w / V
    const uint4 *vmeta_out;      // out-lists (the in-lists themselves when the graph is undirected)
    uint16_t *x[2];              // amounts popped by the running / next sweep, bf16, vertex-major [V][Sr] (pull.cuh)
    int32_t pull_gshift;         // log2 of the lanes that share a vertex in a sweep (pull.cuh, PullGeom)
    unsigned long long dense_enter_edges;  // an iteration expected to traverse at least this many in-edges runs as a sweep
    unsigned long long dense_exit_edges;   // ... and below this the loop goes back to scatter iterations
    int32_t pull_warp_min, pull_big_min, pull_big_chunk;  // out-degree tiers of a sweep; entries per chunk of the grid tier
    int32_t pull_warp_units;     // several sources: 0 = work items go to CTAs, n = to warps, n warp-slices of a tile at a time
    uint32_t pull_tile_mul;      // tile visiting order: tile = (t * mul) mod ntiles, mul coprime to ntiles
    uint32_t *tile_list;         // active tiles of the running dense episode: [3][tile_list_cap], heavy tiles first
    uint32_t tile_list_cap;
    int32_t signed_push;         // variant 0: one pass over both signs instead of the reference's two (see push_persistent)
    double accel_frac;           // Chebyshev-accelerated sweeps while the frontier holds at least this fraction of all (vertex, source)
                                 // pairs; 0 = never (directed windows: the spectrum is not real)
    HubItem *big;                // grid-tier list
    uint32_t bigcap;
    double *bigacc;              // [bigcap][lanes per vertex x sources per lane] partial sums of the grid tier (zero between sweeps)
};

#ifndef DPPR_ITEMS_PER_THREAD
#define DPPR_ITEMS_PER_THREAD 2
#endif
// frontier items a thread pops per tile, at most.  2 rather than 4: the tile arrays shrink from 41 KB to 26 KB per CTA,
// which the SM gives back as L1 (out-degrees and ring slots of popular vertices hit there): -5..-7 % on every probe.
constexpr int kItemsPerThread = DPPR_ITEMS_PER_THREAD;
constexpr int kTileMax = kThreads * kItemsPerThread;    // 512 items per tile

struct PushSmem {
    unsigned long long stage[kStage];
    double t_ru[kTileMax];      // (1-alpha) * claimed residual of the item
    uint32_t t_off[kTileMax];   // exclusive prefix of in-degrees over the tile
    uint32_t t_base[kTileMax];  // ring geometry of the item's in-list
    uint32_t t_head[kTileMax];
    uint32_t t_mask[kTileMax];
    uint32_t t_s[kTileMax];     // source index of the item
    uint32_t h_c0[kHubSmem];
    uint32_t scan[kWarps + 1];
    unsigned int stage_cnt;
    unsigned int gbase;
    unsigned int pl_n, pl_cnt;   // pull.cuh: CTA-tier list length, legal-count accumulator
    unsigned long long pl_edges;
    uint32_t bar_units;          // grid-wide total reported at the last grid barrier
    int abort_flag;
    int dedupe;                  // signed pass: the next frontier is past half the queue -- stamp crossings from now on
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define DPPR_TL(tl, slot) do { if ((tl) && threadIdx.x == 0) (tl)[slot] = global_ns(); } while (0)

// phase 0 / 1: the reference's two passes (gpu/PPRCommon.cuh:6-11); kSignedPhase: both signs at once (variant 0, below)
constexpr int kSignedPhase = 2;
__device__ __forceinline__ bool legal_push(double x, int phase, double eps) {
    return phase == 0 ? (x > eps) : phase == 1 ? (x < -eps) : (fabs(x) > eps);
}

// ---- next-frontier staging ------------------------------------------------------------------
// must be called by all 32 lanes of a warp together
__device__ __forceinline__ void stage_push(bool want, unsigned long long item, PushSmem &sm,
                                           unsigned long long *qout, unsigned int *cnt_out, uint32_t qcap,
                                           PushCtrl *ctrl) {
    const unsigned m = __ballot_sync(kFull, want);
    if (m == 0) return;
    const int n = __popc(m);
    const int leader = __ffs(m) - 1;
    unsigned base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(&sm.stage_cnt, (unsigned)n);
    base = __shfl_sync(kFull, base, leader);
    // stage_cnt only ever grows between flushes (so reserved positions below kStage have exactly one
    // writer and no holes); positions at or beyond kStage spill straight to the global queue
    const unsigned pos = base + __popc(m & lanemask_lt());
    const bool spill = want && pos >= (unsigned)kStage;
    if (want && !spill) sm.stage[pos] = item;
    if (base + n > (unsigned)kStage) {  // warp-uniform
        const unsigned m2 = __ballot_sync(kFull, spill);
        const int leader2 = __ffs(m2) - 1;
        unsigned g = 0;
        if ((int)lane_id() == leader2) g = atomicAdd(cnt_out, (unsigned)__popc(m2));
        g = __shfl_sync(kFull, g, leader2);
        if (spill) {
            const unsigned gp = g + __popc(m2 & lanemask_lt());
            if (gp < qcap) __stcg(&qout[gp], item);
            else atomicOr(&ctrl->errflags, kErrQueue);
        }
    }
}

// CTA-wide; leaves stage empty.  Contains barriers: call from uniform control flow.
__device__ __forceinline__ void stage_flush(PushSmem &sm, unsigned long long *qout, unsigned int *cnt_out,
                                            uint32_t qcap, PushCtrl *ctrl) {
    __syncthreads();
    const unsigned n = sm.stage_cnt < (unsigned)kStage ? sm.stage_cnt : (unsigned)kStage;  // the rest spilled
    if (n) {
        if (threadIdx.x == 0) sm.gbase = atomicAdd(cnt_out, n);
        __syncthreads();
        const unsigned g = sm.gbase;
        for (unsigned i = threadIdx.x; i < n; i += kThreads) {
            if (g + i < qcap) __stcg(&qout[g + i], sm.stage[i]);
            else atomicOr(&ctrl->errflags, kErrQueue);
        }
        __syncthreads();
        if (threadIdx.x == 0) sm.stage_cnt = 0;
    }
    __syncthreads();
}

// ---- traversed in-edges, kEdgeUnroll per thread ------------------------------------------------
// The push is bound by dependent memory round trips (slot -> out-degree -> atomic), not by bandwidth:
// every thread therefore keeps kEdgeUnroll independent chains in flight.  Stage 1 issues all slot
// loads, stage 2 all out-degree loads, stage 3 all FP64 atomics, stage 4 consumes the returned old
// values for the enqueue rule.
// Owner accessors keep per-slot register state down to (neighbour id, owner index): the per-owner
// values (scaled residual, source row base) are re-read from shared memory / uniform registers.
struct TileOwner {
    const PushSmem &sm;
    uint32_t lo[kEdgeUnroll];
    __device__ __forceinline__ double ru_scaled(int k) const { return sm.t_ru[lo[k]]; }
    __device__ __forceinline__ uint32_t s(int k) const { return sm.t_s[lo[k]]; }
};
struct HubOwner {
    double ru_scaled_;
    uint32_t s_;
    __device__ __forceinline__ double ru_scaled(int) const { return ru_scaled_; }
    __device__ __forceinline__ uint32_t s(int) const { return s_; }
};

template <int VAR, class Owner>
__device__ __forceinline__ void push_edges(const PushArgs &a, PushSmem &sm, const Owner &ow,
                                           const uint32_t (&nbr)[kEdgeUnroll], const bool (&active)[kEdgeUnroll],
                                           unsigned long long *qout, unsigned int *cnt_out, int phase, int level) {
    int32_t dv[kEdgeUnroll];
    double add[kEdgeUnroll], old[kEdgeUnroll];
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k) dv[k] = active[k] ? __ldg(&a.outdeg[nbr[k]]) : 0;
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k) {
        old[k] = 0.0; add[k] = 0.0;
        if (active[k]) {
            add[k] = ow.ru_scaled(k) / (double)(dv[k] + 1);  // (1-alpha)*ru/(outdeg(v)+1), gpu/ExpandRev.cuh:71-72
            old[k] = atomicAdd(&a.r[(unsigned long long)nbr[k] * a.Sr + ow.s(k)], add[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < kEdgeUnroll; ++k) {
        bool want = false;
        if (active[k]) {
            const double cur = old[k] + add[k];
            if (VAR == 0 || VAR == 1) {  // threshold crossing, gpu/ExpandRev.cuh:75
                want = !legal_push(old[k], phase, a.eps) && legal_push(cur, phase, a.eps);
                // variant 0's signed pass: mixed-sign adds can carry a residual out of the band, back in and out again within
                // one iteration.  The twin is harmless (its pop finds an exact zero) but takes a queue slot, and nothing bounds
                // the number of re-crossings: once the next frontier is past half the queue, a crossing only enqueues if it is the
                // first one stamped at this level.  (Stamping every crossing costs one more dependent atomic per pop: +17 % on
                // BASELINE configs[1].)  At most one unstamped + one stamped entry per pair: the queue holds twice the pairs.
                if (VAR == 0 && want && phase == kSignedPhase && sm.dedupe)
                    want = atomicExch(&a.status[(unsigned long long)nbr[k] * a.Sr + ow.s(k)], level) < level;
            } else if (legal_push(cur, phase, a.eps)) {  // status stamp, gpu/ExpandRev.cuh:254-257
                want = atomicExch(&a.status[(unsigned long long)nbr[k] * a.Sr + ow.s(k)], level) < level;
            }
        }
        stage_push(want, ((unsigned long long)ow.s(k) << 32) | nbr[k], sm, qout, cnt_out, a.qcap, a.ctrl);
    }
}

// ---- seeds --------------------------------------------------------------------------------------
__device__ void seed_pass(const PushArgs &a, PushSmem &sm, int phase, unsigned long long *qout,
                          unsigned int *cnt_out, bool with_candidates = true, bool with_alt = false) {
    const uint32_t ncand = a.init_mode ? 1u : __ldcg(a.ncand);
    const unsigned long long total = with_candidates ? (unsigned long long)ncand * (unsigned)a.S : 0ull;
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
            const uint32_t u = a.init_mode ? (uint32_t)a.src[s] : a.cand[c];
            const double x = __ldcg(&a.r[(unsigned long long)u * a.Sr + s]);
            want = legal_push(x, phase, a.eps);
            item = ((unsigned long long)s << 32) | u;
            if (want) mx = fmax(mx, fabs(x));
        }
        stage_push(want, item, sm, qout, cnt_out, a.qcap, a.ctrl);
        if ((rd & 3) == 3) stage_flush(sm, qout, cnt_out, a.qcap, a.ctrl);
    }
    stage_flush(sm, qout, cnt_out, a.qcap, a.ctrl);
    if (with_alt && a.qalt) {
        // residuals of this phase's sign that the last dense episode of the previous phase left behind (pull_compact): a
        // sweep pushes both signs, so they may sit anywhere, not only at repaired vertices.  An episode absorbs EVERY
        // residual beyond eps, so after one this list is complete and the candidates are not scanned at all -- no
        // (source, vertex) pair is ever seeded twice (variants 1-3 would pop a duplicate twice).
        const unsigned long long nalt = __ldcg(&a.ctrl->nalt[phase & 1]);
        const unsigned long long rounds2 = (nalt + stride - 1) / stride;
        for (unsigned long long rd = 0; rd < rounds2; ++rd) {
            const unsigned long long j = rd * stride + (unsigned long long)blockIdx.x * kThreads + threadIdx.x;
            bool want = false;
            unsigned long long item = 0;
            if (j < nalt) {
                item = __ldcg(&a.qalt[j]);
                const double x = __ldcg(&a.r[(unsigned long long)(uint32_t)item * a.Sr + (item >> 32)]);
                want = legal_push(x, phase, a.eps);
                if (want) mx = fmax(mx, fabs(x));
            }
            stage_push(want, item, sm, qout, cnt_out, a.qcap, a.ctrl);
            if ((rd & 3) == 3) stage_flush(sm, qout, cnt_out, a.qcap, a.ctrl);
        }
        stage_flush(sm, qout, cnt_out, a.qcap, a.ctrl);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmax(mx, __shfl_xor_sync(kFull, mx, off));
    if (lane_id() == 0 && mx > 0.0) atomicMax(&a.ctrl->theta0[phase & 1], (unsigned long long)__double_as_longlong(mx));
}

// ---- pre pass: variants 1,3 snapshot + zero (gpu/Inspect.cuh:52-65); variant 2 status stamp ------
template <int VAR>
__device__ void pre_pass(const PushArgs &a, const unsigned long long *qin, double *qr, uint32_t n, int level) {
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const unsigned long long item = __ldcg(&qin[i]);
        const unsigned long long idx = (unsigned long long)(uint32_t)item * a.Sr + (item >> 32);
        if (VAR == 1 || VAR == 3) {
            const double x = __ldcg(&a.r[idx]);
            __stcg(&qr[i], x);
            atomicAdd(&a.p[idx], a.alpha * x);  // RED: result unused
            __stcg(&a.r[idx], 0.0);
        } else if (VAR == 2) {
            __stcg(&a.status[idx], level);
        }
    }
}

// ---- post pass (variant 2): r[u] -= ru, still legal -> next frontier (gpu/ExpandRev.cuh:709-743) ---
__device__ void post_pass(const PushArgs &a, PushSmem &sm, const unsigned long long *qin, const double *qr, uint32_t n,
                          unsigned long long *qout, unsigned int *cnt_out, int phase) {
    const uint32_t stride = gridDim.x * kThreads;
    const uint32_t rounds = (n + stride - 1) / stride;
    for (uint32_t rd = 0; rd < rounds; ++rd) {
        const uint32_t i = rd * stride + blockIdx.x * kThreads + threadIdx.x;
        bool want = false;
        unsigned long long item = 0;
        if (i < n) {
            item = __ldcg(&qin[i]);
            const unsigned long long idx = (unsigned long long)(uint32_t)item * a.Sr + (item >> 32);
            const double ru = __ldcg(&qr[i]);
            const double old = atomicAdd(&a.r[idx], -ru);
            want = legal_push(old - ru, phase, a.eps);
        }
        stage_push(want, item, sm, qout, cnt_out, a.qcap, a.ctrl);
        if ((rd & 3) == 3) stage_flush(sm, qout, cnt_out, a.qcap, a.ctrl);
    }
    stage_flush(sm, qout, cnt_out, a.qcap, a.ctrl);
}

// ---- hubs popped in the previous iteration: their edge chunks are dealt round-robin to all CTAs -------
// Producers number the chunks with one packed 64-bit atomic ((hubs << 32) | chunks), so hub k's
// chunk0 is non-decreasing in k and the owner of chunk c is found by binary search.
template <int VAR>
__device__ void expand_hubs(const PushArgs &a, PushSmem &sm, const HubItem *hin, unsigned long long hpk,
                            unsigned long long *qout, unsigned int *cnt_out, int phase, int level,
                            unsigned long long &edges_acc) {
    const uint32_t nh = min((uint32_t)(hpk >> 32), a.hcap), nchunks = (uint32_t)hpk;  // (beyond hcap: dropped, kErrHubQ raised)
    if (nh == 0) return;
    const bool cached = nh <= (uint32_t)kHubSmem;
    if (cached) {
        for (uint32_t h = threadIdx.x; h < nh; h += kThreads) sm.h_c0[h] = __ldcg(&hin[h].chunk0);
        __syncthreads();
    }
    for (uint32_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        if (threadIdx.x == 0) sm.dedupe = __ldcg(cnt_out) > a.qcap / 2;
        __syncthreads();
        uint32_t lo = 0, hi = nh;  // last hub with chunk0 <= c
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            const uint32_t v = cached ? sm.h_c0[mid] : __ldcg(&hin[mid].chunk0);
            if (v <= c) lo = mid; else hi = mid;
        }
        const unsigned long long item = __ldcg(&hin[lo].item);
        const double ru_scaled = (1.0 - a.alpha) * __ldcg(&hin[lo].ru);
        const uint32_t c0 = cached ? sm.h_c0[lo] : __ldcg(&hin[lo].chunk0);
        const uint32_t s = (uint32_t)(item >> 32), v = (uint32_t)item;
        const uint4 m = __ldg(&a.vmeta[v]);
        const uint32_t deg = m.z, mask = m.w - 1u;
        const uint32_t e0 = (c - c0) * (uint32_t)kHubChunk;
        uint32_t nbr[kEdgeUnroll];
        bool active[kEdgeUnroll];
#pragma unroll
        for (int k = 0; k < kEdgeUnroll; ++k) {
            const uint32_t e = e0 + k * kThreads + threadIdx.x;
            active[k] = e < deg;
            nbr[k] = active[k] ? (uint32_t)__ldg(&a.pool[m.x + ((m.y + e) & mask)]) : 0u;
        }
        const HubOwner ow{ru_scaled, s};
        push_edges<VAR>(a, sm, ow, nbr, active, qout, cnt_out, phase, level);
        if (threadIdx.x == 0) edges_acc += min(deg - e0, (uint32_t)kHubChunk);
        stage_flush(sm, qout, cnt_out, a.qcap, a.ctrl);
    }
    __syncthreads();  // h_c0 is reused by the next call
}

// ---- frontier tiles: pop, edge-balanced expansion -----------------------------------------------------
// A tile is the contiguous slice of the frontier one CTA handles at a time: at most kTileMax items,
// and no more than ceil(n / grid) so that one pass of the grid covers the frontier whenever
// n <= grid * kTileMax.  Threads pop up to kItemsPerThread items each with all loads of a stage in
// flight together; the in-degrees are prefix-summed in shared memory and the concatenated edge range
// is walked kEdgeUnroll edges per thread per round.
template <int VAR>
__device__ void expand_tiles(const PushArgs &a, PushSmem &sm, const unsigned long long *qin, double *qr, uint32_t n,
                             unsigned long long *qout, unsigned int *cnt_out, HubItem *hout,
                             unsigned long long *hpk_out, int phase, int level, double theta, unsigned long long &edges_acc,
                             unsigned long long &carried_acc, unsigned long long *tl = nullptr) {
    if (n == 0) return;
    // Every CTA gets the same number k of equal tiles.  k = 1 while one pass of the grid covers the frontier with
    // tiles of at most 512 items (the latency-bound regime: one pop/scan/edge/flush chain per iteration); larger
    // frontiers are cut into tiles of at most `tile_cap` items: later tiles then pop residuals that already contain
    // the earlier tiles' pushes (Gauss-Seidel-ish: -20 % traversals on the Orkut-shaped probe) and the stages of
    // different CTAs interleave instead of marching in lock step (16.4 -> 11.6 ms per batch there).
    const uint32_t per_cta = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t k = per_cta <= 512u ? 1u : (per_cta + (uint32_t)a.tile_cap - 1) / (uint32_t)a.tile_cap;
    uint32_t tile_items = (n + gridDim.x * k - 1) / (gridDim.x * k);
    tile_items = tile_items < 8u ? 8u : (tile_items > (uint32_t)kTileMax ? (uint32_t)kTileMax : tile_items);
    const uint32_t ntiles = (n + tile_items - 1) / tile_items;
    const uint32_t ipt = (tile_items + kThreads - 1) / kThreads;  // 1..kItemsPerThread
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t tbase = tile * tile_items;
        if (threadIdx.x == 0) sm.dedupe = __ldcg(cnt_out) > a.qcap / 2;  // (published by the barrier that follows the pops)
        // ---- pop: stage 1 items, stage 2 ring metadata + residual claim, stage 3 estimate update ----
        unsigned long long item[kItemsPerThread];
        bool have[kItemsPerThread], carry[kItemsPerThread];
#pragma unroll
        for (int k = 0; k < kItemsPerThread; ++k) {
            const uint32_t j = k * kThreads + threadIdx.x;
            have[k] = j < tile_items && tbase + j < n;
            carry[k] = false;
            item[k] = have[k] ? __ldcg(&qin[tbase + j]) : 0ull;
        }
#pragma unroll
        for (int k = 0; k < kItemsPerThread; ++k) {
            const uint32_t j = k * kThreads + threadIdx.x;
            if (j < tile_items) {
                uint32_t deg = 0;
                if (have[k]) {
                    const uint32_t s = (uint32_t)(item[k] >> 32), v = (uint32_t)item[k];
                    const unsigned long long idx = (unsigned long long)v * a.Sr + s;
                    const uint4 m = __ldg(&a.vmeta[v]);
                    double ru;
                    if (VAR == 0) {
                        ru = __longlong_as_double((long long)atomicExch((unsigned long long *)&a.r[idx], 0ull));
                        if (fabs(ru) <= theta && fabs(ru) > a.eps) {
                            // threshold schedule: too small to be worth a push yet -- put it back (RED) and carry the
                            // vertex to the next frontier, where it will have collected more mass
                            atomicAdd(&a.r[idx], ru);
                            carry[k] = true;
                        } else {
                            atomicAdd(&a.p[idx], a.alpha * ru);  // result unused: fire-and-forget RED, no extra round trip
                        }
                    } else if (VAR == 2) {
                        ru = __ldcg(&a.r[idx]);            // live read, kept until the post pass subtracts it
                        __stcg(&qr[tbase + j], ru);
                        atomicAdd(&a.p[idx], a.alpha * ru);
                    } else {
                        ru = __ldcg(&qr[tbase + j]);       // taken by the snapshot pass
                    }
                    deg = (carry[k] || ru == 0.0) ? 0u : m.z;  // (an exact zero: a duplicate whose twin took everything)
                    if (deg >= (uint32_t)a.hub_degree) {
                        const uint32_t nch = (deg + kHubChunk - 1) / kHubChunk;
                        const unsigned long long old = atomicAdd(hpk_out, (1ull << 32) | nch);
                        const uint32_t hp = (uint32_t)(old >> 32);
                        if (hp < a.hcap) {
                            __stcg(&hout[hp].item, item[k]);
                            __stcg(&hout[hp].ru, ru);
                            __stcg(&hout[hp].chunk0, (uint32_t)old);
                        } else {
                            atomicOr(&a.ctrl->errflags, kErrHubQ);
                        }
                        deg = 0;
                    }
                    sm.t_ru[j] = (1.0 - a.alpha) * ru;
                    sm.t_base[j] = m.x;
                    sm.t_head[j] = m.y;
                    sm.t_mask[j] = m.w - 1u;
                    sm.t_s[j] = s;
                }
                sm.t_off[j] = deg;
            }
        }
        if (VAR == 0) {
#pragma unroll
            for (int k = 0; k < kItemsPerThread; ++k) {  // uniform: every lane takes part in the ballots
                stage_push(carry[k], item[k], sm, qout, cnt_out, a.qcap, a.ctrl);
                carried_acc += carry[k] ? 1 : 0;
            }
        }
        __syncthreads();
        if (tile == blockIdx.x) DPPR_TL(tl, 2);
        // ---- exclusive prefix of the degrees (blocked: thread t owns items [t*ipt, t*ipt+ipt)) ----
        uint32_t d[kItemsPerThread], sum = 0;
        const uint32_t jb = threadIdx.x * ipt;
#pragma unroll
        for (int k = 0; k < kItemsPerThread; ++k) {
            d[k] = ((uint32_t)k < ipt && jb + k < tile_items) ? sm.t_off[jb + k] : 0u;
            sum += d[k];
        }
        uint32_t total;
        uint32_t off = block_exclusive_sum<uint32_t>(sum, sm.scan, total);
#pragma unroll
        for (int k = 0; k < kItemsPerThread; ++k) {
            if ((uint32_t)k < ipt && jb + k < tile_items) {
                sm.t_off[jb + k] = off;
                off += d[k];
            }
        }
        __syncthreads();
        if (tile == blockIdx.x) DPPR_TL(tl, 3);
        // ---- edges ----
        for (uint32_t e0 = 0; e0 < total; e0 += kHubChunk) {
            TileOwner ow{sm, {}};
            uint32_t nbr[kEdgeUnroll];
            bool active[kEdgeUnroll];
#pragma unroll
            for (int k = 0; k < kEdgeUnroll; ++k) {
                const uint32_t e = e0 + k * kThreads + threadIdx.x;
                active[k] = e < total;
                nbr[k] = 0;
                if (active[k]) {
                    // owner = last j with t_off[j] <= e   (zero-degree items share an offset with their successor)
                    uint32_t lo = 0, hi = tile_items;
                    while (hi - lo > 1) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (sm.t_off[mid] <= e) lo = mid; else hi = mid;
                    }
                    ow.lo[k] = lo;
                    const uint32_t kk = e - sm.t_off[lo];
                    nbr[k] = (uint32_t)__ldg(&a.pool[sm.t_base[lo] + ((sm.t_head[lo] + kk) & sm.t_mask[lo])]);
                }
            }
            push_edges<VAR>(a, sm, ow, nbr, active, qout, cnt_out, phase, level);
        }
        if (threadIdx.x == 0) edges_acc += total;
        if (tile == blockIdx.x) DPPR_TL(tl, 4);
        stage_flush(sm, qout, cnt_out, a.qcap, a.ctrl);
        if (tile == blockIdx.x) DPPR_TL(tl, 5);
    }
}

// ---- software grid barrier (all CTAs co-resident: cooperative launch) --------------------------------
// bar.sync orders the CTA's earlier global writes before thread 0's release-add (cumulativity); the
// acquire-load that observes the last arrival makes every other CTA's writes visible, and the
// trailing bar.sync passes that on to the rest of the CTA.
__device__ __forceinline__ void bar_arrive_release(unsigned *addr) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(addr) : "memory");
}
__device__ __forceinline__ unsigned bar_load_acquire(const unsigned *addr) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned bar_load_relaxed(const unsigned *addr) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void bar_arrive_release_u64(unsigned long long *addr, unsigned long long inc) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(addr), "l"(inc) : "memory");
}
__device__ __forceinline__ unsigned long long bar_load_acquire_u64(const unsigned long long *addr) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long bar_load_relaxed_u64(const unsigned long long *addr) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
    return v;
}

// The barrier also SUMS one 32-bit number per CTA (`units`, e.g. the in-edges it traversed in this iteration / 64) at
// no extra cost: it rides in the high half of the arrival atomic, and every thread finds the grid-wide total of this
// barrier in sm.bar_units afterwards.  Two words are used alternately, so that the total a slow CTA reads cannot
// already contain contributions to the next barrier (a CTA reaches barrier g+2 only after everyone has left g+1).
struct GridBar {
    unsigned gen = 0;
    uint32_t prev_hi0 = 0u, prev_hi1 = 0u;  // thread 0: high half of each word at its previous use
};
// (Smem: anything with `abort_flag` and `bar_units` members in shared memory)
// SUM = false: plain barrier (nothing reported, sm.bar_units untouched).
template <bool SUM = false, class Smem>
__device__ __forceinline__ bool grid_barrier(PushCtrl *c, GridBar &gb, Smem &sm, uint32_t units = 0) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++gb.gen;
        const unsigned wsel = gb.gen & 1u;
        const unsigned target = ((gb.gen + 1u) >> 1) * gridDim.x;
        unsigned long long *word = &c->bar64[wsel];
        bar_arrive_release_u64(word, SUM ? (((unsigned long long)units << 32) | 1ull) : 1ull);
        const long long t0 = clock64();
        bool ok = true;
        // poll with a relaxed load: ld.acquire makes ptxas emit CCTL.IVALL (whole-L1 invalidate) after EVERY
        // poll, which stalls the poller and flushes the L1 of the CTAs still working on this SM (ncu: 31 % of
        // all stall samples).  One acquire after the last arrival is enough.
        while ((unsigned)bar_load_relaxed_u64(word) < target) {
            if (clock64() - t0 > 8000000000ll) {  // ~4 s: a CTA is missing, give up loudly instead of hanging
                atomicOr(&c->errflags, kErrWatchdog);
                ok = false;
                break;
            }
        }
        const uint32_t hi = (uint32_t)(bar_load_acquire_u64(word) >> 32);
        if (SUM) {
            sm.bar_units = hi - (wsel ? gb.prev_hi1 : gb.prev_hi0);
            if (wsel) gb.prev_hi1 = hi; else gb.prev_hi0 = hi;
        }
        sm.abort_flag = ok ? 0 : 1;
    }
    __syncthreads();
    return sm.abort_flag == 0;
}

}  // namespace dppr
#include "pull.cuh"
namespace dppr {

// ---- the persistent kernel ----------------------------------------------------------------------------
// DENSE != 0: with the switch to gather sweeps (variant 0 only).  Separate instantiations, so that the scatter-only
// kernels keep their register allocation.
#ifndef DPPR_DENSE1_MIN_BLOCKS
#define DPPR_DENSE1_MIN_BLOCKS 3
#endif
#ifndef DPPR_DENSE8_MIN_BLOCKS
#define DPPR_DENSE8_MIN_BLOCKS 3
#endif
// DENSE = sources a lane of a sweep takes: 0 (scatter only), 1 (one source) or 8 (several).  The switching kernels run
// 3 CTAs per SM (80 registers): scripts/micro/gather_mlp.cu shows that the random-row bandwidth of an SM grows with its
// resident WARPS, not with the loads a warp keeps in flight, while round 1's 64-register build spilled (5.2 G local loads
// per refresh on BASELINE configs[3]); 3 x 8 warps with 4 / 8 gathers per lane measured best (profiles/README.md).
template <int VAR, int DENSE>
__global__ void __launch_bounds__(kThreads, DENSE == 8 ? DPPR_DENSE8_MIN_BLOCKS : DENSE == 1 ? DPPR_DENSE1_MIN_BLOCKS : DPPR_MIN_BLOCKS)
    push_persistent(const __grid_constant__ PushArgs a) {
    __shared__ PushSmem sm;
    PushCtrl *c = a.ctrl;
    if (threadIdx.x == 0) { sm.stage_cnt = 0; sm.abort_flag = 0; }
    __syncthreads();
    GridBar gen;
    unsigned long long edges_acc = 0, pops_acc = 0, hubs_acc = 0, carried_acc = 0, gath_acc = 0;
    uint32_t sweeps_done = 0;
    float rate_reg = DENSE ? fmaxf(__ldcg(&c->rate_ns[0]), __ldcg(&c->rate_ns[1])) : 0.f;  // block 0 / thread 0: running estimate
    const int level0 = __ldcg(&c->level);
    uint32_t it = 0, iters_done = 0;  // `it` indexes the rotating slots (skips one value per phase change)
    // Phases alternate in sign: 0 pushes residuals above eps, 1 those below -eps (the reference runs exactly these two,
    // gpu/PPRGPU.cuh:128-163).  A dense episode pushes BOTH signs at once and may leave residuals of the other sign behind
    // (qalt): they seed the next phase, and further phases run while an episode keeps leaving some (normally none do).
    //
    // Variant 0 pushes both signs in ONE pass (a.signed_push): its frontier is deduped by the threshold crossing
    // |old| <= eps < |old + add| and popped by an exchange that claims the whole residual, so mixed-sign adds are harmless;
    // a residual that leaves the band, is pulled back and leaves again within one iteration is kept from being enqueued
    // twice by a per-level stamp.  The two passes of the reference each have their own ramp-up and tail of tiny iterations, and residuals of
    // opposite sign cancel instead of being pushed separately: about half the iterations on the L2-resident configs, where an
    // iteration costs a fixed ~6 us of dependent round trips.  Variants 1-3 keep the reference's two passes.
    const bool signed_mode = VAR == 0 && a.signed_push && !a.init_mode;
    const int nphases = (a.init_mode || signed_mode) ? 1 : 2;
    bool alive = true;
    bool prev_episode = false;  // an episode ran in the previous phase (every CTA takes the same decisions)
    for (int phase_i = 0; phase_i < 8 && alive; ++phase_i) {
        const int phase = signed_mode ? kSignedPhase : (phase_i & 1);
        if (phase_i >= nphases && !(DENSE && !signed_mode && prev_episode && __ldcg(&c->nalt[phase & 1]) != 0u)) break;  // (uniform: this word was last
        // written before the barrier that ended the previous phase; it is reset below, after the next barrier, and
        // appended to only by episodes of the NEXT phase)
        if (phase_i > 0) {
            // Phase change.  Slow CTAs may still be polling cnt[it % 3] (== 0) to leave the loop below,
            // so the new seeds must not land in that slot: skip one iteration index.  The slots the
            // skipped iteration would have cleared are cleared here -- their last readers passed the
            // barrier that ended iteration it-1, their next writers run after the barrier below.
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                c->cnt[(it + 2) % 3] = 0;
                c->hpk[(it + 1) % 3] = 0;
            }
            ++it;
        }
        // seeds: the repaired vertices -- or, if an episode ran in the previous phase, what it left of this phase's sign
        seed_pass(a, sm, phase, a.q[it & 1], &c->cnt[it % 3], phase_i < nphases && !prev_episode, DENSE && prev_episode && !signed_mode);
        prev_episode = false;  // (from here on: "an episode ran in THIS phase")
        if (!(alive = grid_barrier(c, gen, sm))) break;
        if (DENSE && blockIdx.x == 0 && threadIdx.x == 0) c->nalt[phase & 1] = 0u;  // consumed (episodes of this phase fill nalt[phase ^ 1])
        const bool carrying = VAR == 0 && a.carry_gamma > 0.0 && a.carry_gamma < 1.0;
        double theta = carrying ? __longlong_as_double((long long)__ldcg(&c->theta0[phase & 1])) * a.carry_scale : a.eps;
        uint32_t n_prev = 0;
        bool fresh_phase = true;
        double t_prev = 0.0;  // in-edges the previous scatter iteration traversed, grid-wide (DENSE)
        while (alive) {
            bool want_dense = false;
            unsigned long long dense_hpk = 0;
            float dense_rate = 0.f;
            while (true) {
                const uint32_t n = min(__ldcg(&c->cnt[it % 3]), a.qcap);  // (beyond qcap the writers dropped items and raised kErrQueue)
                const unsigned long long hpk = __ldcg(&c->hpk[(it + 2) % 3]);
                // (slot it & 1 was written two iterations ago: the other slot may be being rewritten right now)
                const float rate = DENSE ? __ldcg(&c->rate_ns[it & 1]) : 0.f, sw = DENSE ? __ldcg(&c->sweep_ns) : 0.f;
                const uint32_t nh = (uint32_t)(hpk >> 32);
                if (n == 0 && nh == 0) break;
                if ((int)it >= a.max_iters) {
                    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&c->errflags, kErrWatchdog);
                    alive = false;
                    break;
                }
                if (DENSE && !carrying && fresh_phase && rate > 0.f && sw > 0.f) {
                    // first iteration of a phase: no history -- the seeds are batch endpoints, take
                    // twice the average in-degree for them
                    const double pred = 2.0 * (double)n * (double)a.avg_indeg;
                    if (pred * (double)rate > 1.25 * (double)sw || pred >= 0.5 * (double)a.qcap) {
                        want_dense = true;
                        dense_hpk = hpk;
                        dense_rate = rate;
                        break;
                    }
                }
                if (DENSE && !carrying && n_prev != 0) {
                    // expected work of this iteration: the tiles the previous one traversed, scaled by the frontier
                    // growth, plus the hub chunks it left for this one.  Once both costs have been measured on this
                    // engine (a sweep costs the same whatever the frontier; a scatter iteration pays per edge) the
                    // measured figures decide, before that the static estimate of the host.
                    const double pred = t_prev * ((double)n / (double)n_prev) + 0.75 * (double)kHubChunk * (double)(uint32_t)hpk;
                    // (a traversed edge yields at most one new frontier item: never let the predicted next frontier come
                    // near the queue capacity -- with many sources V x S exceeds it, the sweeps need no queue)
                    const bool enter = ((rate > 0.f && sw > 0.f) ? pred * (double)rate > 1.25 * (double)sw
                                                                  : pred >= (double)a.dense_enter_edges) ||
                                       pred >= 0.5 * (double)a.qcap;
                    if (enter) {  // handled by the outer loop
                        want_dense = true;
                        dense_hpk = hpk;
                        dense_rate = rate;
                        break;
                    }
                }
                n_prev = n;
                fresh_phase = false;
                const unsigned long long edges_before = edges_acc;
                const unsigned long long t_iter0 = (DENSE && blockIdx.x == 0 && threadIdx.x == 0) ? global_ns() : 0ull;
                if (blockIdx.x == 0 && threadIdx.x == 0) {  // slots nobody reads or writes during this iteration
                    c->cnt[(it + 2) % 3] = 0;
                    c->hpk[(it + 1) % 3] = 0;
                    hubs_acc += nh;
                    pops_acc += n;
                    if (a.iterlog && (int)iters_done < a.iterlog_cap) {
                        unsigned long long t;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                        a.iterlog[iters_done] = make_uint4(n, (uint32_t)hpk, (uint32_t)t, (uint32_t)(t >> 32));
                    }
                }
                const int level = level0 + (int)it + 1;
                unsigned long long *tl = (a.ctalog && (int)iters_done == a.probe_iter) ? a.ctalog + (size_t)blockIdx.x * 8 : nullptr;
                DPPR_TL(tl, 0);
                const unsigned long long *qin = a.q[it & 1];
                unsigned long long *qout = a.q[(it + 1) & 1];
                if (VAR != 0) {
                    pre_pass<VAR>(a, qin, a.qr[it & 1], n, level);
                    if (!(alive = grid_barrier(c, gen, sm))) break;
                }
                expand_hubs<VAR>(a, sm, a.hub[(it + 1) & 1], hpk, qout, &c->cnt[(it + 1) % 3], phase, level, edges_acc);
                DPPR_TL(tl, 1);
                expand_tiles<VAR>(a, sm, qin, a.qr[it & 1], n, qout, &c->cnt[(it + 1) % 3], a.hub[it & 1],
                                  &c->hpk[it % 3], phase, level, fmax(theta, a.eps), edges_acc, carried_acc, tl);
                theta *= a.carry_gamma;
                DPPR_TL(tl, 6);
                if (VAR == 2) {
                    if (!(alive = grid_barrier(c, gen, sm))) break;
                    post_pass(a, sm, qin, a.qr[it & 1], n, qout, &c->cnt[(it + 1) % 3], phase);
                }
                if (!(alive = grid_barrier<(DENSE != 0)>(c, gen, sm, DENSE ? (uint32_t)((edges_acc - edges_before) >> 6) : 0u))) break;
                if (DENSE) {
                    const double t_all = 64.0 * (double)sm.bar_units;  // tiles + hub chunks, grid-wide
                    t_prev = fmax(0.0, t_all - 0.75 * (double)kHubChunk * (double)(uint32_t)hpk);
                    if (blockIdx.x == 0 && threadIdx.x == 0) {
                        if (t_all >= 262144.0) {
                            const float sample = (float)((double)(global_ns() - t_iter0) / t_all);
                            rate_reg = rate_reg > 0.f ? 0.5f * (rate_reg + sample) : sample;
                        }
                        c->rate_ns[it & 1] = rate_reg;
                    }
                }
                DPPR_TL(tl, 7);
                ++it;
                ++iters_done;
            }
            if (!DENSE || !want_dense || !alive) break;
            // the frontier is large: this and the following iterations run as gather sweeps (pull.cuh)
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                c->cnt[(it + 2) % 3] = 0;
                c->hpk[(it + 1) % 3] = 0;
                hubs_acc += (uint32_t)(dense_hpk >> 32);
            }
            prev_episode = true;
            n_prev = 0;  // (no growth estimate for the first scatter iteration after the sweeps)
            fresh_phase = false;
            t_prev = 0.0;
            DenseIO io{edges_acc, gath_acc, pops_acc, iters_done, sweeps_done, gen, dense_rate};
            alive = dense_mode<(DENSE ? DENSE : 1)>(a, sm, c, phase, it, dense_hpk, io);
            edges_acc = io.edges_acc; gath_acc = io.gath; pops_acc = io.pops_acc;
            iters_done = io.iters_done; sweeps_done = io.sweeps_done; gen = io.gen;
            ++it;
        }
    }
    // carried_acc is per thread: fold it over the CTA with one atomic per warp
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        carried_acc += __shfl_xor_sync(kFull, carried_acc, off);
        gath_acc += __shfl_xor_sync(kFull, gath_acc, off);
    }
    if (lane_id() == 0 && carried_acc) atomicAdd(&c->carried, carried_acc);
    if (lane_id() == 0 && gath_acc) atomicAdd(&c->gath, gath_acc);
    if (threadIdx.x == 0) {
        if (edges_acc) atomicAdd(&c->edges, edges_acc);
        if (blockIdx.x == 0) {
            c->sweeps = sweeps_done;
            c->iters = iters_done;
            c->pops = pops_acc;
            c->hubs = hubs_acc;
            c->level = level0 + (int)it + 2;
        }
    }
}

// ---- stepwise mode: the same device functions, one launch per sub-pass ----------------------------------
__global__ void __launch_bounds__(kThreads) push_step_seed(const PushArgs a, uint32_t it, int phase) {
    __shared__ PushSmem sm;
    if (threadIdx.x == 0) { sm.stage_cnt = 0; sm.abort_flag = 0; }
    __syncthreads();
    seed_pass(a, sm, phase, a.q[it & 1], &a.ctrl->cnt[it % 3]);
}

template <int VAR>
__global__ void __launch_bounds__(kThreads) push_step_pre(const PushArgs a, uint32_t it, int level) {
    pre_pass<VAR>(a, a.q[it & 1], a.qr[it & 1], a.ctrl->cnt[it % 3], level);
}

template <int VAR>
__global__ void __launch_bounds__(kThreads) push_step_expand(const PushArgs a, uint32_t it, int phase, int level, double theta) {
    __shared__ PushSmem sm;
    if (threadIdx.x == 0) { sm.stage_cnt = 0; sm.abort_flag = 0; }
    __syncthreads();
    PushCtrl *c = a.ctrl;
    const uint32_t n = min(c->cnt[it % 3], a.qcap);
    const unsigned long long hpk = c->hpk[(it + 2) % 3];
    const uint32_t nh = (uint32_t)(hpk >> 32);
    unsigned long long edges_acc = 0, carried_acc = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c->cnt[(it + 2) % 3] = 0;
        c->hpk[(it + 1) % 3] = 0;
        c->pops += n;
        c->hubs += nh;
        c->iters += 1;
    }
    expand_hubs<VAR>(a, sm, a.hub[(it + 1) & 1], hpk, a.q[(it + 1) & 1], &c->cnt[(it + 1) % 3], phase, level, edges_acc);
    expand_tiles<VAR>(a, sm, a.q[it & 1], a.qr[it & 1], n, a.q[(it + 1) & 1], &c->cnt[(it + 1) % 3], a.hub[it & 1],
                      &c->hpk[it % 3], phase, level, theta, edges_acc, carried_acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) carried_acc += __shfl_xor_sync(kFull, carried_acc, off);
    if (lane_id() == 0 && carried_acc) atomicAdd(&c->carried, carried_acc);
    if (threadIdx.x == 0 && edges_acc) atomicAdd(&c->edges, edges_acc);
}

__global__ void __launch_bounds__(kThreads) push_step_post(const PushArgs a, uint32_t it, int phase) {
    __shared__ PushSmem sm;
    if (threadIdx.x == 0) { sm.stage_cnt = 0; sm.abort_flag = 0; }
    __syncthreads();
    post_pass(a, sm, a.q[it & 1], a.qr[it & 1], a.ctrl->cnt[it % 3], a.q[(it + 1) & 1], &a.ctrl->cnt[(it + 1) % 3], phase);
}

// ---- state initialisation (gpu/PPRCommon.cuh:13-22) ------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    state_init(double *__restrict__ p, double *__restrict__ r, int32_t *__restrict__ status, int32_t V, int64_t Sr, int S,
               const int32_t *__restrict__ src) {
    const int64_t total = (int64_t)V * Sr;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int64_t v = i / Sr;
        const int s = (int)(i - v * Sr);
        p[i] = 0.0;
        r[i] = (s < S && v == src[s]) ? 1.0 : 0.0;
        if (status) status[i] = -1;
    }
}

}  // namespace dppr
