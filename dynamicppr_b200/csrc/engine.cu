// engine.cu -- see engine.cuh.  Reference call stack being replaced: SURVEY.md 3.1 / 3.2.
#include "engine.cuh"
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>

namespace dppr {

namespace {

// `st`: where the refresh's counters are -- the control block itself, or the sum over source panels
__global__ void gather_record(BatchRecord *rec, const PushCtrl *ctrl, const PushCtrl *st, const uint32_t *counters,
                              const unsigned long long *pool_top) {
    BatchRecord r{};
    r.iters = st->iters; r.pops = st->pops; r.edges = st->edges; r.gath = st->gath;
    r.hubs = st->hubs; r.carried = st->carried; r.dpops = st->dpops;
    r.walk_slots = st->walk_slots; r.walk_pairs = st->walk_pairs; r.units = st->units;
    r.pool_top = pool_top[0]; r.pool_leaked = pool_top[1];
    r.sweeps = st->sweeps;
    r.nseg_in = counters[0];
    r.nseg_out = counters[1];
    r.njobs = counters[2] + counters[5];
    r.errflags = ctrl->errflags;
    r.arrived = 1;
    *rec = r;
}

// a panel's refresh is done: add its counters to the running sums
__global__ void accumulate_ctrl(PushCtrl *acc, const PushCtrl *ctrl) {
    acc->iters += ctrl->iters; acc->pops += ctrl->pops; acc->edges += ctrl->edges; acc->gath += ctrl->gath;
    acc->hubs += ctrl->hubs; acc->carried += ctrl->carried; acc->dpops += ctrl->dpops;
    acc->walk_slots += ctrl->walk_slots; acc->walk_pairs += ctrl->walk_pairs; acc->units += ctrl->units;
    acc->sweeps += ctrl->sweeps;
}

__global__ void fold_window_errors(PushCtrl *ctrl, int *win_err) {
    if (*win_err) atomicOr(&ctrl->errflags, *win_err);
}

// caller ids -> internal ids for the source list (one launch instead of one blocking 4-byte copy per source)
__global__ void translate_sources(const int32_t *__restrict__ src_in, const uint32_t *__restrict__ perm, int32_t *__restrict__ src_out, int S) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) src_out[i] = (int32_t)perm[src_in[i]];
}

// the environment is a development aid for A/B runs: consulted ONCE, in the constructor, and only for knobs the
// caller left at 0 in dppr_config::tuning
int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}
double env_real(const char *name, double dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atof(v) : dflt;
}
int pick_int(int32_t field, const char *env, int dflt) { return field != 0 ? (int)field : env_int(env, dflt); }
double pick_real(double field, const char *env, double dflt) { return field != 0.0 ? field : env_real(env, dflt); }

Tuning resolve_tuning(const dppr_tuning &t) {
    Tuning r;
    r.relabel = t.relabel != 0 ? t.relabel > 0 : env_int("DPPR_RELABEL", 1) != 0;
    r.relabel_blocks = std::max(1, pick_int(t.relabel_blocks, "DPPR_RELABEL_BLOCKS", 1024));
    r.relabel_both = pick_int(t.relabel_both, "DPPR_RELABEL_BOTH", 0) > 0;
    r.ctas_per_sm = std::max(1, pick_int(t.ctas_per_sm, "DPPR_CTAS_PER_SM", 4));
    r.tile_cap = std::min(std::max(pick_int(t.tile_cap, "DPPR_TILE_CAP", 128), 8), kTileMax);
    r.max_iters = std::max(1, pick_int(t.max_iters, "DPPR_MAX_ITERS", 400000));
    r.dense = pick_int(t.dense, "DPPR_DENSE", 0);
    r.dense_div = pick_real(t.dense_div, "DPPR_DENSE_DIV", 4.0);
    r.dense_min_edges = pick_real(t.dense_min_edges, "DPPR_DENSE_MIN_EDGES", -1.0);  // (< 0: by variant, see the constructor)
    if (std::getenv("DPPR_DENSE_DIV") && t.dense_div == 0.0 && std::atof(std::getenv("DPPR_DENSE_DIV")) <= 0.0) r.dense = -1;  // round-1 spelling of "off"
    if (std::getenv("DPPR_DENSE_MIN_EDGES") && t.dense_min_edges == 0.0 && std::atof(std::getenv("DPPR_DENSE_MIN_EDGES")) <= 0.0) r.dense_min_edges = 0.0;
    r.pull_group = std::min(std::max(pick_int(t.pull_group, "DPPR_PULL_GROUP", 16), 1), 16);
    r.pull_warp_min = std::max(1, pick_int(t.pull_warp_min, "DPPR_PULL_WARP_MIN", 32));
    r.pull_big_min = pick_int(t.pull_big_min, "DPPR_PULL_BIG_MIN", 0);
    r.pull_big_chunk = pick_int(t.pull_big_chunk, "DPPR_PULL_BIG_CHUNK", 0);
    r.carry_gamma = pick_real(t.carry_gamma, "DPPR_CARRY_GAMMA", 1.0);
    r.carry_scale = pick_real(t.carry_scale, "DPPR_CARRY_SCALE", 0.01);
    r.dense_accel = pick_int(t.dense_accel, "DPPR_DENSE_ACCEL", 0);
    r.signed_push = pick_int(t.signed_push, "DPPR_SIGNED_PUSH", 0);
    r.accel_frac = env_real("DPPR_ACCEL_FRAC", 0.9);
    r.window_path = pick_int(t.window_path, "DPPR_WINDOW_PATH", 0);
    if (t.window_path == 0 && !env_int("DPPR_COOP_WINDOW", 1)) r.window_path = 1;   // round-1 spellings
    else if (t.window_path == 0 && !env_int("DPPR_FUSED_WINDOW", 1)) r.window_path = 2;
    r.iterlog = pick_int(t.iterlog, "DPPR_ITERLOG", 0) > 0;
    r.probe_iter = pick_int(t.probe_iter, "DPPR_PROBE_ITER", 10);
    r.panel_sources = std::max(1, pick_int(t.panel_sources, "DPPR_PANEL_SOURCES", 128));
    r.pull_warp_units = std::max(0, pick_int(t.pull_warp_units, "DPPR_PULL_WARP_UNITS", 8));  // (-1 -> 0: CTA items)
    return r;
}

template <int VAR>
void *persistent_kernel_v(int dense) {
    return dense == 8 ? (void *)push_persistent<VAR, 8> : dense == 1 ? (void *)push_persistent<VAR, 1> : (void *)push_persistent<VAR, 0>;
}
// the persistent kernel of a variant: scatter only (dense = 0), or with the switch to gather sweeps for one (1) / several (8) sources
void *persistent_kernel(int variant, int dense) {
    switch (variant) {
        case 0: return persistent_kernel_v<0>(dense);
        case 1: return persistent_kernel_v<1>(dense);
        case 2: return persistent_kernel_v<2>(dense);
        default: return persistent_kernel_v<3>(dense);
    }
}

}  // namespace

int Engine::grid_for(int64_t n) const {
    int64_t g = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sm_count_ * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

Engine::Engine(const dppr_config &cfg) : cfg_(cfg) {
    if (cfg.vertex_count <= 0) throw InvalidArgument("vertex_count must be positive");
    if (cfg.window_edges <= 0) throw InvalidArgument("window_edges must be positive");
    if (cfg.max_batch_edges < 0 || cfg.max_batch_edges > cfg.window_edges)
        throw InvalidArgument("max_batch_edges must be in [0, window_edges]");
    if (cfg.variant < DPPR_OPTIMIZED || cfg.variant > DPPR_VANILLA)
        throw InvalidArgument("variant must be 0..3 (Meta.h:11-17)");
    if (cfg.n_sources < 1 || cfg.sources == nullptr) throw InvalidArgument("at least one source vertex is required");
    if (cfg.engine_mode != DPPR_ENGINE_AUTO && cfg.engine_mode != DPPR_ENGINE_STEPWISE && cfg.engine_mode != DPPR_ENGINE_LEVELSYNC)
        throw InvalidArgument("engine_mode must be one of DPPR_ENGINE_{AUTO,STEPWISE,LEVELSYNC}");
    mode_ = cfg.engine_mode == DPPR_ENGINE_AUTO ? DPPR_ENGINE_LEVELSYNC : cfg.engine_mode;
    if (const char *force = std::getenv("DPPR_FORCE_ENGINE")) {  // A/B runs without touching the caller (read once, here)
        const int f = std::atoi(force);
        if (f == DPPR_ENGINE_STEPWISE || f == DPPR_ENGINE_LEVELSYNC) mode_ = f;
    }
    tn_ = resolve_tuning(cfg.tuning);
    debug_ = env_int("DPPR_DEBUG", 0) != 0;  // (like every DPPR_* variable: read here, once)
    if (cfg_.alpha <= 0.0) cfg_.alpha = 0.15;
    if (cfg_.alpha >= 1.0) throw InvalidArgument("alpha must be in (0, 1)");
    if (cfg_.epsilon <= 0.0) cfg_.epsilon = 1e-9;
    if (cfg_.pool_factor <= 0.0) cfg_.pool_factor = 8.0;
    if (cfg_.hub_degree <= 0) cfg_.hub_degree = env_int("DPPR_HUB_DEGREE", 64);
    V_ = cfg.vertex_count;
    D_ = cfg.directed ? 1 : 2;
    W_ = cfg.window_edges;
    Ew_ = W_ * D_;
    Bmax_ = cfg.max_batch_edges;
    Nb_ = std::max<int64_t>(2 * D_ * Bmax_, 1);
    S_ = cfg.n_sources;
    n_panels_ = div_up(S_, tn_.panel_sources);
    Pw_ = div_up(S_, n_panels_);  // (equal panels: 1000 sources -> 8 x 125, not 7 x 128 + 104)
    Sr_ = Pw_ == 1 ? 1 : ((int64_t)Pw_ + 7) / 8 * 8;
    panel_stride_ = (int64_t)cfg.vertex_count * Sr_;
    sources_.assign(cfg.sources, cfg.sources + S_);
    cfg_.sources = sources_.data();
    for (int32_t s : sources_)
        if (s < 0 || s >= V_) throw InvalidArgument("source vertex id out of range");
    if ((uint64_t)V_ >= (1ull << 31)) throw InvalidArgument("vertex ids must fit in 31 bits");
    if (Ew_ >= (int64_t)0xffffffffll) throw InvalidArgument("window too large for 32-bit CSR offsets");
    key_bits_ = bits_for((uint64_t)(V_ - 1));
    {
        // dense iterations (pull.cuh): variant 0 of the level-synchronous engine; an iteration runs as a gather sweep
        // once it is expected to traverse at least (E_w + 2 V) x sources / dense_div in-edges.
        // Small windows stay with the scatter-only kernel: their iterations are bound by the chain of dependent round
        // trips, which a sweep does not shorten (BASELINE configs[1]: 43 vs 18 us), and the kernel that can switch
        // carries more loop state, which costs its scatter iterations +9..+24 % when they are latency-bound (LJ/4,
        // youtube) and nothing when they are bandwidth-bound (Twitter-shaped).
        // (every variant: the four differ in how the SCATTER form reads residuals and dedupes its frontier; a gather sweep decides each
        // (vertex, source) once, so the episodes are the same for all of them -- round-1 verdict, item 7)
        // Variants 1-3 always get the kernel that can switch: their scatter levels cost 2-3 passes each, so sweeps pay off
        // on much smaller windows (BASELINE configs[2], LiveJournal-shaped, 6.9 M edges: 7.7 / 5.9 / 8.4 -> 3.2 ms per batch)
        // and where they do not (configs[1]) the device-side cost model simply never enters them (2.19 vs 2.24 ms).
        const bool can = mode_ == DPPR_ENGINE_LEVELSYNC && tn_.dense_div > 0.0;
        const double min_edges = tn_.dense_min_edges >= 0.0 ? tn_.dense_min_edges : (cfg.variant == DPPR_OPTIMIZED ? 2.0e7 : 0.0);
        dense_ = can && tn_.dense >= 0 && (tn_.dense > 0 || (double)cfg.window_edges * (cfg.directed ? 1 : 2) * Pw_ >= min_edges);
        outlists_ = dense_ && D_ == 1;
        // several sources: a lane takes 8 sources (16 bytes of an x row), G = 2^gshift adjacent lanes share a vertex (pull.cuh)
        pull_gshift_ = 0;
        if (Pw_ > 1) {
            const int chunks = (int)(Sr_ / 8);
            while ((1 << pull_gshift_) < std::min(chunks, tn_.pull_group) && pull_gshift_ < 4) ++pull_gshift_;  // (<= 16 lanes: 128 accumulator columns)
        }
        // out-lists of big_min or more entries are cut into chunks any warp of the grid takes: a warp streams a list at
        // kPullUnroll rows per memory round trip, so the wider the rows (the fewer vertices a warp holds) the shorter the chunks
        pull_big_min_ = tn_.pull_big_min > 0 ? tn_.pull_big_min : (Pw_ == 1 ? 4096 : 1024);
        pull_big_min_ = std::max(pull_big_min_, tn_.pull_warp_min);
        pull_big_chunk_ = tn_.pull_big_chunk > 0 ? tn_.pull_big_chunk : std::max(32, pull_big_min_ / 4);
    }

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw CudaFailure(std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (cfg.device < 0 || cfg.device >= ndev) throw InvalidArgument("device ordinal out of range");
    dev_ = cfg.device;
    DPPR_CUDA(cudaSetDevice(dev_));
    cudaDeviceProp prop;
    DPPR_CUDA(cudaGetDeviceProperties(&prop, dev_));
    sm_count_ = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) throw CudaFailure("device does not support cooperative launch");
    DPPR_CUDA(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));

    // cooperative grid per variant: every CTA must be co-resident for the software grid barrier
    const int dense_kind = dense_ ? (Pw_ == 1 ? 1 : 8) : 0;
    void *kern[4] = {persistent_kernel(0, cfg_.variant == 0 ? dense_kind : 0), persistent_kernel(1, cfg_.variant == 1 ? dense_kind : 0),
                     persistent_kernel(2, cfg_.variant == 2 ? dense_kind : 0), persistent_kernel(3, cfg_.variant == 3 ? dense_kind : 0)};
    // the switching kernels keep the relaxation factor of an accelerated sweep in (dynamic) shared memory (pull.cuh)
    dyn_smem_ = dense_ ? (size_t)env_int("DPPR_DYN_SMEM", 16) : 0;
    if (dyn_smem_ < 16 && dense_) dyn_smem_ = 16;
    if (const char *cv = std::getenv("DPPR_CARVEOUT")) DPPR_CUDA(cudaFuncSetAttribute(kern[cfg_.variant], cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(cv)));
    if (dyn_smem_) DPPR_CUDA(cudaFuncSetAttribute(kern[cfg_.variant], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem_));
    for (int v = 0; v < 4; ++v) {
        int per_sm = 0;
        DPPR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern[v], kThreads, v == cfg_.variant ? dyn_smem_ : 0));
        if (per_sm < 1) throw CudaFailure("push kernel does not fit on an SM");
        coop_grid_[v] = std::min(per_sm, tn_.ctas_per_sm) * sm_count_;
        if (debug_) std::fprintf(stderr, "[dppr] variant %d: %d CTAs per SM fit, grid %d\n", v, per_sm, coop_grid_[v]);
    }
    {
        int per_sm = 0;
        DPPR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (void *)win_update_coop, kThreads, 0));
        coop_win_grid_ = per_sm >= 1 ? std::min(kCoopMaxTiles, per_sm * sm_count_) : 0;
    }
    // window
    log_.alloc((size_t)W_);
    vmeta_.alloc((size_t)V_);
    outdeg_.alloc((size_t)V_);
    if (outlists_) vmeta_out_.alloc((size_t)V_);
    double pc = cfg_.pool_factor * (double)Ew_ * (outlists_ ? 2.0 : 1.0) + 4096.0;
    if (pc > 4294967295.0) pc = 4294967295.0;
    pool_cap_ = (unsigned long long)pc;
    pool_.alloc((size_t)pool_cap_);
    pool_top_.alloc(2);
    DPPR_CUDA(cudaMemsetAsync(pool_top_.ptr, 0, pool_top_.bytes(), st_));
    {
        // free stacks: class c (ranges of 2^c slots) gets room for min(pool >> c, max(65536, V >> (c - 2))) ranges -- the
        // number of rings of a size class falls geometrically on a power-law graph; 8 bytes per vertex in total
        std::vector<uint32_t> off(33, 0);
        for (int c = 0; c < 32; ++c) {
            unsigned long long room = 0;
            if (c >= 2) room = std::min<unsigned long long>(pool_cap_ >> c, std::max<unsigned long long>(65536ull, (unsigned long long)V_ >> (c - 2)));
            off[c + 1] = off[c] + (uint32_t)room;
        }
        fcount_.alloc(32);
        foff_.alloc(33);
        fstack_.alloc(std::max<size_t>(off[32], 1));
        npend_.alloc(2);
        for (int i = 0; i < 2; ++i) pend_[i].alloc((size_t)Nb_ * (outlists_ ? 2 : 1));
        DPPR_CUDA(cudaMemsetAsync(fcount_.ptr, 0, fcount_.bytes(), st_));
        DPPR_CUDA(cudaMemsetAsync(npend_.ptr, 0, npend_.bytes(), st_));
        DPPR_CUDA(cudaMemcpyAsync(foff_.ptr, off.data(), sizeof(uint32_t) * 33, cudaMemcpyHostToDevice, st_));
        DPPR_CUDA(cudaStreamSynchronize(st_));  // `off` is a local
    }
    // batch scratch
    arriving_.alloc((size_t)std::max<int64_t>(Bmax_, 1));
    for (int i = 0; i < 2; ++i) {
        akey_[i].alloc((size_t)Nb_);
        aval_[i].alloc((size_t)Nb_);
        if (D_ == 1) {
            bkey_[i].alloc((size_t)Nb_);
            bval_[i].alloc((size_t)Nb_);
        }
    }
    sort_scratch_.alloc(sort_scratch_elems(Nb_) + scan_scratch_elems(Nb_));
    flags_.alloc((size_t)Nb_);
    segA_vertex_.alloc((size_t)Nb_); segA_start_.alloc((size_t)Nb_ + 1); segA_first_.alloc((size_t)Nb_); segA_of_.alloc((size_t)Nb_);
    counters_.alloc(8);
    tile_heads_.alloc(kCoopMaxTiles);
    segA_ = Segments{segA_vertex_.ptr, segA_start_.ptr, segA_first_.ptr, segA_of_.ptr, counters_.ptr + 0};
    if (D_ == 1) {
        segB_vertex_.alloc((size_t)Nb_); segB_start_.alloc((size_t)Nb_ + 1); segB_first_.alloc((size_t)Nb_); segB_of_.alloc((size_t)Nb_);
        segB_ = Segments{segB_vertex_.ptr, segB_start_.ptr, segB_first_.ptr, segB_of_.ptr, counters_.ptr + 1};
    } else {
        segB_ = segA_;
    }
    ins_pos_.alloc((size_t)Nb_);
    jobs_.alloc((size_t)Nb_);
    if (outlists_) { ins_posB_.alloc((size_t)Nb_); jobsB_.alloc((size_t)Nb_); }
    if (dense_) {
        for (int i = 0; i < 2; ++i) {
            x_[i].alloc((size_t)n_panels_ * panel_stride_ + 8);
            DPPR_CUDA(cudaMemsetAsync(x_[i].ptr, 0, x_[i].bytes(), st_));
        }
        const int lanes_sources = Pw_ == 1 ? 1 : 8 << pull_gshift_;                // sources one pass over a vertex covers
        const int n_cg = (int)((Sr_ + lanes_sources - 1) / lanes_sources);          // chunk groups
        tile_cap_ = (uint32_t)div_up(V_, kThreads >> pull_gshift_) * (uint32_t)n_cg;
        tile_list_.alloc((size_t)tile_cap_ * 3);
        bigcap_ = (uint32_t)std::min<int64_t>((Ew_ / pull_big_min_ + 64) * n_cg, 1 << 24);
        big_.alloc(bigcap_);
        bigacc_.alloc((size_t)bigcap_ * lanes_sources);
        DPPR_CUDA(cudaMemsetAsync(bigacc_.ptr, 0, bigacc_.bytes(), st_));
    }
    seg_d0_.alloc((size_t)Nb_);
    delta_.alloc((size_t)Nb_ * Sr_);
    DPPR_CUDA(cudaMemsetAsync(delta_.ptr, 0, delta_.bytes(), st_));
    DPPR_CUDA(cudaMemsetAsync(counters_.ptr, 0, counters_.bytes(), st_));
    // state
    p_.alloc((size_t)n_panels_ * panel_stride_);
    r_.alloc((size_t)n_panels_ * panel_stride_);
    if (n_panels_ > 1) ctrl_acc_.alloc(1);
    if (dense_ && Pw_ > 1) {  // popped-amount sums of a sweep episode (pull.cuh): zero whenever no episode is running
        pacc_.alloc((size_t)n_panels_ * panel_stride_);
        DPPR_CUDA(cudaMemsetAsync(pacc_.ptr, 0, pacc_.bytes(), st_));
    }
    // level stamps: the frontier dedupe of variants 2, 3 -- and of variant 0's signed pass (push.cuh)
    if (cfg_.variant >= DPPR_EAGER || (cfg_.variant == DPPR_OPTIMIZED && tn_.signed_push >= 0)) status_.alloc((size_t)n_panels_ * panel_stride_);
    src_.alloc((size_t)S_);
    DPPR_CUDA(cudaMemcpyAsync(src_.ptr, sources_.data(), sizeof(int32_t) * S_, cudaMemcpyHostToDevice, st_));
    // push queues: a frontier holds each (source, vertex) at most once -- twice in variant 0's signed pass (push_edges), plus
    // what the CTAs hold staged while the stamping starts
    const bool signed0 = cfg_.variant == DPPR_OPTIMIZED && tn_.signed_push >= 0;
    int64_t qc = cfg_.frontier_capacity > 0 ? cfg_.frontier_capacity
                                            : std::min<int64_t>((int64_t)V_ * Pw_ * (signed0 ? 2 : 1) + (signed0 ? (1 << 20) : 0), (int64_t)1 << 29);
    qc = std::max<int64_t>(qc, 1024);
    if (qc > 0xfffffff0ll) qc = 0xfffffff0ll;
    qcap_ = (uint32_t)qc;
    // hub list: one entry per popped (source, vertex) of in-degree >= hub_degree.  With the switching kernel the large
    // frontiers run as sweeps (hubs still pending at the switch are un-popped), so a quarter of the worst case is plenty;
    // overflow is detected (DPPR_DEVERR_HUBQ), never silent
    int64_t hc = std::max<int64_t>(1024, (Ew_ / cfg_.hub_degree + 1) * Pw_);
    if (dense_ && Pw_ > 8) hc = std::max<int64_t>(1 << 20, hc / 4);
    hcap_ = (uint32_t)std::min<int64_t>(qc, hc);
    for (int i = 0; i < 2; ++i) {
        q_[i].alloc(qcap_);
        if (cfg_.variant != DPPR_OPTIMIZED) qr_[i].alloc(qcap_);
        hub_[i].alloc(hcap_);
    }
    if (dense_) qalt_.alloc(qcap_);
    ctrl_.alloc(1);
    DPPR_CUDA(cudaMemsetAsync(ctrl_.ptr, 0, sizeof(PushCtrl), st_));
    dev_record_.alloc(1);
    if (tn_.iterlog) {
        iterlog_.alloc(kIterLogCap);
        ctalog_.alloc((size_t)8 * 148 * 16);
        DPPR_CUDA(cudaMemsetAsync(ctalog_.ptr, 0, ctalog_.bytes(), st_));
    }
    for (int i = 0; i < kStageSlots; ++i) {
        hstage_[i].alloc((size_t)std::max<int64_t>(Bmax_, 1));
        DPPR_CUDA(cudaEventCreateWithFlags(&hstage_free_[i], cudaEventDisableTiming));
    }
    DPPR_CUDA(cudaStreamSynchronize(st_));
}

Engine::~Engine() {
    cudaSetDevice(dev_);
    if (st_) cudaStreamSynchronize(st_);
    for (auto &es : events_)
        for (auto &ev : es.ev)
            if (ev) cudaEventDestroy(ev);
    for (auto &ev : hstage_free_)
        if (ev) cudaEventDestroy(ev);
    if (st_) cudaStreamDestroy(st_);
}

PoolFree Engine::pool_free_view(int parity) {
    return PoolFree{fcount_.ptr, foff_.ptr, fstack_.ptr, pend_[parity].ptr, npend_.ptr + parity, pool_top_.ptr + 1};
}

// ---- per-batch bookkeeping ------------------------------------------------------------------------------------
// Phase timings use a fixed ring of event sets (no event is created in the streaming loop after the first
// kEventSlots batches); the set of batch k - kEventSlots is resolved into plain floats before batch k reuses it.
void Engine::resolve_events(EventSet &es) {
    if (es.batch < 0 || es.batch >= (int64_t)meta_.size()) return;
    BatchMeta &m = meta_[(size_t)es.batch];
    if (m.resolved) return;
    for (int i = 4; i >= 0; --i)
        if (es.used[i]) { DPPR_CUDA(cudaEventSynchronize(es.ev[i])); break; }
    auto ms = [&](int a, int b) -> float {
        if (!es.used[a] || !es.used[b]) return 0.f;
        float t = 0.f;
        return cudaEventElapsedTime(&t, es.ev[a], es.ev[b]) == cudaSuccess ? t : 0.f;
    };
    m.ms[0] = m.has_upload ? ms(0, 1) : 0.f;
    m.ms[1] = m.has_window ? ms(1, 2) : 0.f;
    m.ms[2] = m.has_window ? ms(2, 3) : 0.f;
    m.ms[3] = ms(3, 4);
    m.resolved = true;
}

void Engine::begin_batch(int64_t edges, int64_t entries) {
    check_health();
    meta_.emplace_back();
    BatchMeta &m = meta_.back();
    m.edges = edges;
    m.entries = entries;
    const int64_t k = (int64_t)meta_.size() - 1;
    BatchRecord *rec = record_slot((size_t)k);
    std::memset(rec, 0, sizeof(BatchRecord));
    if (cfg_.record_timing) {
        EventSet &es = events_[k % kEventSlots];
        resolve_events(es);
        es.batch = k;
        for (bool &u : es.used) u = false;
    }
}

void Engine::record(int which) {
    if (!cfg_.record_timing) return;
    EventSet &es = events_[((int64_t)meta_.size() - 1) % kEventSlots];
    if (!es.ev[which]) DPPR_CUDA(cudaEventCreate(&es.ev[which]));
    DPPR_CUDA(cudaEventRecord(es.ev[which], st_));
    es.used[which] = true;
}

BatchRecord *Engine::record_slot(size_t k) {
    while (records_.size() * kRecordsPerChunk <= k) {
        records_.emplace_back(new PinnedBuf<BatchRecord>());
        records_.back()->alloc(kRecordsPerChunk);
        std::memset(records_.back()->ptr, 0, sizeof(BatchRecord) * kRecordsPerChunk);
    }
    return records_[k / kRecordsPerChunk]->ptr + (k % kRecordsPerChunk);
}

void Engine::finish_record() {
    fold_window_errors<<<1, 1, 0, st_>>>(ctrl_.ptr, (int *)(counters_.ptr + 3)); ++launch_counter();
    gather_record<<<1, 1, 0, st_>>>(dev_record_.ptr, ctrl_.ptr, n_panels_ > 1 ? ctrl_acc_.ptr : ctrl_.ptr, counters_.ptr, pool_top_.ptr); ++launch_counter();
    DPPR_CUDA(cudaGetLastError());
    DPPR_CUDA(cudaMemcpyAsync(record_slot(meta_.size() - 1), dev_record_.ptr, sizeof(BatchRecord),
                              cudaMemcpyDeviceToHost, st_));
}

// Device-side failures (pool / queue exhaustion, watchdog, FIFO underflow, bad ids) travel to the host in the batch
// record.  Every record that has landed is inspected here -- at the start of each batch, and by the synchronising calls --
// and the first flagged one fails the engine for good: the window graph is no longer the reference's (round-1 verdict:
// inserts were dropped while the call still returned DPPR_OK).
void Engine::check_health() {
    if (failed_code_ == 0) {
        const size_t n = batch_pending_ ? meta_.size() - 1 : meta_.size();
        while (health_checked_ < n) {
            const BatchRecord *rec = record_slot(health_checked_);
            if (!*(volatile const int *)&rec->arrived) break;  // (copy not landed yet: look again at the next call)
            const int f = rec->errflags;
            if (f) {
                std::string what;
                if (f & kErrPool) what += " adjacency pool exhausted (raise dppr_config.pool_factor);";
                if (f & kErrQueue) what += " frontier queue overflow (raise dppr_config.frontier_capacity);";
                if (f & kErrHubQ) what += " hub list overflow (raise dppr_config.frontier_capacity / hub_degree);";
                if (f & kErrWatchdog) what += " push watchdog fired (max_iters / grid barrier);";
                if (f & kErrUnderflow) what += " expiry of an edge the window does not hold;";
                if (f & kErrBadId) what += " edge endpoint outside [0, vertex_count);";
                failed_msg_ = "batch " + std::to_string(health_checked_) + " left device error flags " + std::to_string(f) + ":" + what +
                              " the engine state is no longer valid";
                failed_code_ = (f & (kErrPool | kErrQueue | kErrHubQ)) ? DPPR_E_CAPACITY : (f & kErrBadId) ? DPPR_E_INVALID : DPPR_E_STATE;
                break;
            }
            ++health_checked_;
        }
    }
    if (failed_code_ == DPPR_E_CAPACITY) throw CapacityError(failed_msg_);
    if (failed_code_ == DPPR_E_INVALID) throw InvalidArgument(failed_msg_);
    if (failed_code_) throw StateError(failed_msg_);
}

// ---------------------------------------------------------------------------------------------
// initial window
// ---------------------------------------------------------------------------------------------
// host input is checked before it is staged: a bad id fails the call, nothing reaches the device
// (the reference asserts the same while reading the file, GraphVec.h:55-56)
void Engine::validate_host_ids(const int32_t *pairs, const int32_t *e1, const int32_t *e2, int64_t n) const {
    const uint32_t V = (uint32_t)V_;
    uint32_t worst = 0;
    if (pairs) {
        for (int64_t i = 0; i < 2 * n; ++i) worst = std::max(worst, (uint32_t)pairs[i]);
    } else {
        for (int64_t i = 0; i < n; ++i) worst = std::max(worst, std::max((uint32_t)e1[i], (uint32_t)e2[i]));
    }
    if (worst >= V) throw InvalidArgument("edge endpoint outside [0, vertex_count) (GraphVec.h:55-56 asserts the same)");
}

int2 *Engine::stage_pairs(const int32_t *pairs, const int32_t *e1, const int32_t *e2, int64_t n) {
    const int slot = hstage_next_;
    hstage_next_ = (hstage_next_ + 1) % kStageSlots;
    DPPR_CUDA(cudaEventSynchronize(hstage_free_[slot]));  // the H2D copy that last used this slot is done
    int2 *h = hstage_[slot].ptr;
    if (pairs) {
        std::memcpy(h, pairs, sizeof(int2) * (size_t)n);
    } else {  // EdgeBatch is SoA (EdgeBatch.h:24-26)
        for (int64_t i = 0; i < n; ++i) h[i] = make_int2(e1[i], e2[i]);
    }
    DPPR_CUDA(cudaMemcpyAsync(arriving_.ptr, h, sizeof(int2) * (size_t)n, cudaMemcpyHostToDevice, st_));
    DPPR_CUDA(cudaEventRecord(hstage_free_[slot], st_));
    return arriving_.ptr;
}

void Engine::init_window_soa(const int32_t *e1, const int32_t *e2, int64_t n) {
    if (!e1 || !e2) throw InvalidArgument("null edge arrays");
    if (n != W_) throw InvalidArgument("init_window needs exactly window_edges edges (InitWindowStream asserts the same)");
    std::vector<int32_t> pairs((size_t)n * 2);
    for (int64_t i = 0; i < n; ++i) {
        pairs[2 * i] = e1[i];
        pairs[2 * i + 1] = e2[i];
    }
    init_window_pairs(pairs.data(), n);
}

void Engine::init_window_pairs(const int32_t *pairs, int64_t n) {
    if (!pairs) throw InvalidArgument("null edge array");
    if (n != W_) throw InvalidArgument("init_window needs exactly window_edges edges (InitWindowStream asserts the same)");
    validate_host_ids(pairs, nullptr, nullptr, n);
    DPPR_CUDA(cudaSetDevice(dev_));
    DPPR_CUDA(cudaMemcpyAsync(log_.ptr, pairs, sizeof(int2) * (size_t)W_, cudaMemcpyHostToDevice, st_));
    build_initial_window();
}

void Engine::init_window_device_pairs(const int32_t *dpairs, int64_t n) {
    if (!dpairs) throw InvalidArgument("null device edge array");
    if (n != W_) throw InvalidArgument("init_window needs exactly window_edges edges (InitWindowStream asserts the same)");
    DPPR_CUDA(cudaSetDevice(dev_));
    DPPR_CUDA(cudaMemcpyAsync(log_.ptr, dpairs, sizeof(int2) * (size_t)W_, cudaMemcpyDeviceToDevice, st_));
    build_initial_window();
}

void Engine::build_initial_window() {
    log_start_ = 0;

    DevBuf<uint32_t> key[2], val[2], indeg, caps, rowptr, capbase, scratch, total;
    for (int i = 0; i < 2; ++i) {
        key[i].alloc((size_t)Ew_);
        val[i].alloc((size_t)Ew_);
    }
    indeg.alloc((size_t)V_); caps.alloc((size_t)V_); rowptr.alloc((size_t)V_); capbase.alloc((size_t)V_);
    scratch.alloc(std::max(sort_scratch_elems(Ew_), scan_scratch_elems(V_)));
    total.alloc(1);
    DPPR_CUDA(cudaMemsetAsync(indeg.ptr, 0, indeg.bytes(), st_));
    DPPR_CUDA(cudaMemsetAsync(outdeg_.ptr, 0, outdeg_.bytes(), st_));
    DPPR_CUDA(cudaMemsetAsync(counters_.ptr, 0, counters_.bytes(), st_));
    DPPR_CUDA(cudaMemsetAsync(fcount_.ptr, 0, fcount_.bytes(), st_));
    DPPR_CUDA(cudaMemsetAsync(npend_.ptr, 0, npend_.bytes(), st_));
    DPPR_CUDA(cudaMemsetAsync(pool_top_.ptr, 0, pool_top_.bytes(), st_));
    int *werr = (int *)(counters_.ptr + 3);

    if (tn_.relabel) {
        // internal order = descending out-degree of the initial window (window.cuh, "internal vertex order")
        DevBuf<uint32_t> deg, rk[2], rv[2], rscratch;
        deg.alloc((size_t)V_);
        for (int i = 0; i < 2; ++i) { rk[i].alloc((size_t)V_); rv[i].alloc((size_t)V_); }
        rscratch.alloc(sort_scratch_elems(V_));
        perm_.alloc((size_t)V_); inv_.alloc((size_t)V_);
        DPPR_CUDA(cudaMemsetAsync(deg.ptr, 0, deg.bytes(), st_));
        // key: out-degree (how often the scatter form hits r[w]); with dense iterations on a directed graph out- plus
        // in-degree (the gather form reads x[u] once per IN-edge of u): DPPR_RELABEL_BOTH=1; no difference measured on R-MAT
        const int both = tn_.relabel_both ? 1 : 0;
        relabel_degrees<<<grid_for(W_), kThreads, 0, st_>>>(log_.ptr, W_, D_ == 1, V_, deg.ptr, werr, both); ++launch_counter();
        const int dbits = bits_for((uint64_t)2 * Ew_);
        const uint32_t degmax = (uint32_t)((1ull << dbits) - 1);
        relabel_keys<<<grid_for(V_), kThreads, 0, st_>>>(deg.ptr, degmax, rk[0].ptr, rv[0].ptr, V_); ++launch_counter();
        const int rr = sort_pairs(rk[0].ptr, rv[0].ptr, rk[1].ptr, rv[1].ptr, V_, dbits, rscratch.ptr, st_);
        const uint32_t P = (uint32_t)std::max(1, std::min(tn_.relabel_blocks, V_));
        relabel_assign<<<grid_for(V_), kThreads, 0, st_>>>(rv[rr].ptr, perm_.ptr, inv_.ptr, V_, P); ++launch_counter();
        relabel_log<<<grid_for(W_), kThreads, 0, st_>>>(log_.ptr, W_, perm_.ptr, V_); ++launch_counter();
        // the sources, in internal ids
        DevBuf<int32_t> src_in;
        src_in.alloc((size_t)S_);
        DPPR_CUDA(cudaMemcpyAsync(src_in.ptr, sources_.data(), sizeof(int32_t) * (size_t)S_, cudaMemcpyHostToDevice, st_));
        translate_sources<<<std::max(1, std::min(div_up(S_, 256), 64)), 256, 0, st_>>>(src_in.ptr, perm_.ptr, src_.ptr, S_); ++launch_counter();
        DPPR_CUDA(cudaStreamSynchronize(st_));  // (src_in and the sort buffers above are locals)
    }

    // in-lists, then (directed graphs with dense iterations enabled) out-lists behind them in the same pool
    unsigned long long top = 0;
    for (int side = 0; side < (outlists_ ? 2 : 1); ++side) {
        if (side) DPPR_CUDA(cudaMemsetAsync(indeg.ptr, 0, indeg.bytes(), st_));
        win_init_entries<<<grid_for(W_), kThreads, 0, st_>>>(log_.ptr, W_, D_ == 1, V_, key[0].ptr, val[0].ptr, indeg.ptr,
                                                            outdeg_.ptr, werr, side); ++launch_counter();
        win_init_caps<<<grid_for(V_), kThreads, 0, st_>>>(indeg.ptr, caps.ptr, V_); ++launch_counter();
        exclusive_scan<uint32_t>(indeg.ptr, rowptr.ptr, V_, scratch.ptr, nullptr, st_);
        exclusive_scan<uint32_t>(caps.ptr, capbase.ptr, V_, scratch.ptr, total.ptr, st_);
        uint32_t htotal = 0;
        int herr = 0;
        DPPR_CUDA(cudaMemcpyAsync(&htotal, total.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st_));
        DPPR_CUDA(cudaMemcpyAsync(&herr, werr, sizeof(int), cudaMemcpyDeviceToHost, st_));
        DPPR_CUDA(cudaStreamSynchronize(st_));
        if (herr & kErrBadId) throw InvalidArgument("edge endpoint outside [0, vertex_count) (GraphVec.h:55-56 asserts the same)");
        if (top + htotal > pool_cap_) throw CapacityError("adjacency pool too small for the initial window; raise pool_factor");
        const int res = sort_pairs(key[0].ptr, val[0].ptr, key[1].ptr, val[1].ptr, Ew_, key_bits_, scratch.ptr, st_);
        win_init_fill<<<grid_for(Ew_), kThreads, 0, st_>>>(key[res].ptr, val[res].ptr, Ew_, rowptr.ptr, capbase.ptr, pool_.ptr,
                                                          (uint32_t)top); ++launch_counter();
        win_init_meta<<<grid_for(V_), kThreads, 0, st_>>>(indeg.ptr, caps.ptr, capbase.ptr, side ? vmeta_out_.ptr : vmeta_.ptr, V_,
                                                          (uint32_t)top); ++launch_counter();
        DPPR_CUDA(cudaGetLastError());
        top += htotal;
    }
    DPPR_CUDA(cudaMemcpyAsync(pool_top_.ptr, &top, sizeof(top), cudaMemcpyHostToDevice, st_));
    DPPR_CUDA(cudaStreamSynchronize(st_));
    win_batches_ = 0;
    window_ready_ = true;
    solved_ = false;
    batch_pending_ = false;
}

// ---------------------------------------------------------------------------------------------
// push launch
// ---------------------------------------------------------------------------------------------
void Engine::launch_push(bool init_mode) {
    if (n_panels_ == 1) { launch_push_panel(init_mode, 0); return; }
    DPPR_CUDA(cudaMemsetAsync(ctrl_acc_.ptr, 0, sizeof(PushCtrl), st_));
    for (int k = 0; k < n_panels_; ++k) {
        launch_push_panel(init_mode, k);
        accumulate_ctrl<<<1, 1, 0, st_>>>(ctrl_acc_.ptr, ctrl_.ptr); ++launch_counter();
    }
}

// the refresh of one source panel: its [V][Sr_] block of the state is the whole state as far as the kernels know
void Engine::launch_push_panel(bool init_mode, int panel) {
    DPPR_CUDA(cudaMemsetAsync(ctrl_.ptr, 0, kCtrlZeroBytes, st_));
    const size_t pb = (size_t)panel * (size_t)panel_stride_;
    const int Sk = panel_sources(panel);
    PushArgs a{};
    a.vmeta = vmeta_.ptr; a.pool = pool_.ptr; a.outdeg = outdeg_.ptr;
    a.p = p_.ptr + pb; a.r = r_.ptr + pb; a.status = status_.ptr ? status_.ptr + pb : nullptr;
    a.pacc = pacc_.ptr ? pacc_.ptr + pb : nullptr;
    a.Sr = Sr_; a.S = Sk; a.src = src_.ptr + (size_t)panel * Pw_;
    for (int i = 0; i < 2; ++i) { a.q[i] = q_[i].ptr; a.qr[i] = qr_[i].ptr; a.hub[i] = hub_[i].ptr; }
    a.qalt = qalt_.ptr;
    a.qcap = qcap_; a.hcap = hcap_;
    a.cand = segB_.vertex; a.ncand = segB_.count;
    a.ctrl = ctrl_.ptr;
    a.eps = cfg_.epsilon; a.alpha = cfg_.alpha;
    a.hub_degree = cfg_.hub_degree;
    a.init_mode = init_mode ? 1 : 0;
    a.max_iters = tn_.max_iters;
    // threshold schedule: off by default (gamma 1.0) -- on the L2-resident BASELINE configs the extra, latency-bound
    // iterations cost more than the saved traversals (profiles/README.md); tuning.carry_gamma = 0.7 enables it
    a.carry_gamma = tn_.carry_gamma;
    a.carry_scale = tn_.carry_scale;
    a.tile_cap = tn_.tile_cap;
    a.V = V_;
    a.avg_indeg = (float)((double)Ew_ / (double)V_);
    a.vmeta_out = outlists_ ? vmeta_out_.ptr : vmeta_.ptr;
    a.x[0] = x_[0].ptr ? x_[0].ptr + pb : nullptr; a.x[1] = x_[1].ptr ? x_[1].ptr + pb : nullptr;
    a.pull_gshift = pull_gshift_;
    // cost model: a sweep reads every out-list entry and every vertex row once, whatever the frontier; a scatter
    // iteration pays one random atomic per traversed in-edge.  Measured ratio ~ DPPR_DENSE_DIV (3): Twitter-shaped
    // 3.3 ms per sweep vs 17 edges/ns scattered; Orkut/4 97 us vs 40 edges/ns.
    a.dense_enter_edges = ~0ull;
    if (dense_) a.dense_enter_edges = (unsigned long long)std::max(1.0, ((double)Ew_ + 2.0 * (double)V_) * (double)Sk / tn_.dense_div);
    a.dense_exit_edges = a.dense_enter_edges / 2;
    a.pull_warp_min = tn_.pull_warp_min; a.pull_big_min = pull_big_min_; a.pull_big_chunk = pull_big_chunk_;
    a.pull_warp_units = tn_.pull_warp_units;
    a.big = big_.ptr; a.bigcap = bigcap_; a.bigacc = bigacc_.ptr; a.tile_list = tile_list_.ptr; a.tile_list_cap = tile_cap_;
    a.accel_frac = (D_ == 2 && tn_.dense_accel >= 0) ? tn_.accel_frac : 0.0;
    a.signed_push = tn_.signed_push >= 0 ? 1 : 0;
    {
        const int lanes_sources = Pw_ == 1 ? 1 : 8 << pull_gshift_;
        const uint64_t ntiles = (uint64_t)div_up(V_, kThreads >> pull_gshift_) * (uint64_t)((Sr_ + lanes_sources - 1) / lanes_sources);
        auto gcd = [](uint64_t x, uint64_t y) { while (y) { const uint64_t t = x % y; x = y; y = t; } return x; };
        uint64_t k = std::max<uint64_t>(1, (uint64_t)(0.6180339887 * (double)ntiles)) | 1ull;
        while (gcd(k, ntiles) != 1) k += 2;
        a.pull_tile_mul = (uint32_t)(k % std::max<uint64_t>(ntiles, 1));
        if (a.pull_tile_mul == 0) a.pull_tile_mul = 1;
    }
    a.iterlog = iterlog_.ptr;
    a.iterlog_cap = iterlog_.ptr ? kIterLogCap : 0;
    a.ctalog = ctalog_.ptr;
    a.probe_iter = tn_.probe_iter;
    if (mode_ == DPPR_ENGINE_STEPWISE) {
        launch_push_stepwise(a);
        return;
    }
    void *params[] = {(void *)&a};
    void *kern = persistent_kernel(cfg_.variant, dense_ ? (Pw_ == 1 ? 1 : 8) : 0);
    DPPR_CUDA(cudaLaunchCooperativeKernel(kern, dim3(coop_grid_[cfg_.variant]), dim3(kThreads), params, dyn_smem_, st_));
    ++launch_counter();
}

// Debug / profiling mode with the reference's structure: one launch per sub-pass and a blocking
// read of the frontier counters per iteration (gpu/PPRRevPushGPU.cuh:106-108).
void Engine::launch_push_stepwise(PushArgs &a) {
    const int grid = coop_grid_[cfg_.variant];
    const int var = cfg_.variant;
    uint32_t it = 0;
    PushCtrl h{};
    const bool signed_mode = var == 0 && a.signed_push && !a.init_mode;  // (push_persistent: variant 0 pushes both signs in one pass)
    const int nphases = (a.init_mode || signed_mode) ? 1 : 2;
    for (int phase_i = 0; phase_i < nphases; ++phase_i) {
        const int phase = signed_mode ? kSignedPhase : phase_i;
        if (phase_i > 0) {  // same slot hygiene as push_persistent at a phase change
            DPPR_CUDA(cudaMemsetAsync(&ctrl_.ptr->cnt[(it + 2) % 3], 0, sizeof(unsigned), st_));
            DPPR_CUDA(cudaMemsetAsync(&ctrl_.ptr->hpk[(it + 1) % 3], 0, sizeof(unsigned long long), st_));
            ++it;
        }
        push_step_seed<<<grid, kThreads, 0, st_>>>(a, it, phase); ++launch_counter();
        double theta = -1.0;
        while (true) {
            DPPR_CUDA(cudaMemcpyAsync(&h, ctrl_.ptr, sizeof(PushCtrl), cudaMemcpyDeviceToHost, st_));
            DPPR_CUDA(cudaStreamSynchronize(st_));
            if (h.cnt[it % 3] == 0 && h.hpk[(it + 2) % 3] == 0) break;
            if ((int)it >= a.max_iters) throw CapacityError("push did not converge within DPPR_MAX_ITERS iterations");
            const int level = step_level_ + (int)it + 1;
            if (theta < 0.0) {  // first iteration of the phase: the seeds' largest residual is known now
                double t0;
                std::memcpy(&t0, &h.theta0[phase & 1], sizeof(double));
                const bool carrying = var == 0 && a.carry_gamma > 0.0 && a.carry_gamma < 1.0;
                theta = carrying ? t0 * a.carry_scale : a.eps;
            }
            const double th = std::max(theta, a.eps);
            theta *= a.carry_gamma;
            switch (var) {
                case 0: push_step_expand<0><<<grid, kThreads, 0, st_>>>(a, it, phase, level, th); ++launch_counter(); break;
                case 1:
                    push_step_pre<1><<<grid, kThreads, 0, st_>>>(a, it, level); ++launch_counter();
                    push_step_expand<1><<<grid, kThreads, 0, st_>>>(a, it, phase, level, th); ++launch_counter();
                    break;
                case 2:
                    push_step_pre<2><<<grid, kThreads, 0, st_>>>(a, it, level); ++launch_counter();
                    push_step_expand<2><<<grid, kThreads, 0, st_>>>(a, it, phase, level, th); ++launch_counter();
                    push_step_post<<<grid, kThreads, 0, st_>>>(a, it, phase); ++launch_counter();
                    break;
                default:
                    push_step_pre<3><<<grid, kThreads, 0, st_>>>(a, it, level); ++launch_counter();
                    push_step_expand<3><<<grid, kThreads, 0, st_>>>(a, it, phase, level, th); ++launch_counter();
                    break;
            }
            DPPR_CUDA(cudaGetLastError());
            ++it;
        }
    }
    step_level_ += (int)it + 2;
}

void Engine::solve_initial() {
    if (!window_ready_) throw StateError("dppr_solve_initial before dppr_init_window");
    DPPR_CUDA(cudaSetDevice(dev_));
    DPPR_CUDA(cudaStreamSynchronize(st_));
    meta_.clear();
    for (auto &es : events_) es.batch = -1;
    health_checked_ = 0;
    batch_pending_ = false;
    begin_batch(0, 0);
    record(0);
    for (int k = 0; k < n_panels_; ++k) {
        const size_t pb = (size_t)k * (size_t)panel_stride_;
        state_init<<<grid_for(panel_stride_), kThreads, 0, st_>>>(p_.ptr + pb, r_.ptr + pb, status_.ptr ? status_.ptr + pb : nullptr, V_, Sr_,
                                                                 panel_sources(k), src_.ptr + (size_t)k * Pw_); ++launch_counter();
    }
    DPPR_CUDA(cudaMemsetAsync(ctrl_.ptr, 0, sizeof(PushCtrl), st_));
    DPPR_CUDA(cudaMemsetAsync(counters_.ptr, 0, counters_.bytes(), st_));
    step_level_ = 0;
    record(3);
    launch_push(true);
    record(4);
    finish_record();
    DPPR_CUDA(cudaGetLastError());
    solved_ = true;
}

// ---------------------------------------------------------------------------------------------
// one batch: window update
// ---------------------------------------------------------------------------------------------
void Engine::apply_batch_host_pairs(const int32_t *pairs, int64_t B) {
    if (!pairs && B > 0) throw InvalidArgument("null edge array");
    if (B <= 0 || B > Bmax_) throw InvalidArgument("batch size must be in [1, max_batch_edges]");
    if (!solved_) throw StateError("dppr_apply_batch before dppr_solve_initial");
    if (batch_pending_) throw StateError("dppr_apply_batch twice without dppr_refresh");
    validate_host_ids(pairs, nullptr, nullptr, B);
    DPPR_CUDA(cudaSetDevice(dev_));
    begin_batch(B, 2 * D_ * B);
    record(0);
    int2 *d = stage_pairs(pairs, nullptr, nullptr, B);
    cur().has_upload = true;
    record(1);
    apply_batch_common(d, B);
}

void Engine::apply_batch_host_soa(const int32_t *e1, const int32_t *e2, int64_t B) {
    if ((!e1 || !e2) && B > 0) throw InvalidArgument("null edge arrays");
    if (B <= 0 || B > Bmax_) throw InvalidArgument("batch size must be in [1, max_batch_edges]");
    if (!solved_) throw StateError("dppr_apply_batch before dppr_solve_initial");
    if (batch_pending_) throw StateError("dppr_apply_batch twice without dppr_refresh");
    validate_host_ids(nullptr, e1, e2, B);
    DPPR_CUDA(cudaSetDevice(dev_));
    begin_batch(B, 2 * D_ * B);
    record(0);
    int2 *d = stage_pairs(nullptr, e1, e2, B);
    cur().has_upload = true;
    record(1);
    apply_batch_common(d, B);
}

void Engine::apply_batch_device_pairs(const int32_t *dpairs, int64_t B) {
    if (!dpairs && B > 0) throw InvalidArgument("null device edge array");
    if (B <= 0 || B > Bmax_) throw InvalidArgument("batch size must be in [1, max_batch_edges]");
    if (!solved_) throw StateError("dppr_apply_batch before dppr_solve_initial");
    if (batch_pending_) throw StateError("dppr_apply_batch twice without dppr_refresh");
    DPPR_CUDA(cudaSetDevice(dev_));
    begin_batch(B, 2 * D_ * B);
    record(0);
    record(1);
    apply_batch_common((const int2 *)dpairs, B);
}

void Engine::apply_batch_common(const int2 *arriving, int64_t B) {
    BatchMeta &m = cur();
    const int64_t nA = 2 * D_ * B;  // entries in a group (directed: 2B per group, undirected: 4B in the one group)
    m.has_window = true;
    int *werr = (int *)(counters_.ptr + 3);
    DPPR_CUDA(cudaMemsetAsync(counters_.ptr, 0, sizeof(uint32_t) * 3, st_));
    // ranges released by this batch are collected in pend_[parity]; the first stage of this update hands the previous
    // batch's list (other parity) to the free stacks (window.cuh, PoolFree)
    const int parity = (int)(win_batches_ & 1);
    ++win_batches_;
    DPPR_CUDA(cudaMemsetAsync(npend_.ptr + parity, 0, sizeof(uint32_t), st_));
    const PoolFree fr = pool_free_view(parity);
    const uint2 *pend_prev = pend_[parity ^ 1].ptr;
    const uint32_t *npend_prev = npend_.ptr + (parity ^ 1);
    WindowView wv{V_, vmeta_.ptr, pool_.ptr, outdeg_.ptr, pool_top_.ptr, pool_cap_, werr, fr};
    WindowView wvo{V_, vmeta_out_.ptr, pool_.ptr, outdeg_.ptr, pool_top_.ptr, pool_cap_, werr, fr};  // vmeta null = no out-lists
    if (nA <= kFusedMaxEntries && tn_.window_path == 0) {
        // small batch: the whole update in one single-CTA launch (window_fused.cuh)
        FusedArgs f{};
        f.log = log_.ptr; f.W = W_; f.log_start = log_start_; f.arriving = arriving; f.B = B;
        f.directed = D_ == 1; f.key_bits = key_bits_;
        for (int i = 0; i < 2; ++i) { f.akey[i] = akey_[i].ptr; f.aval[i] = aval_[i].ptr; f.bkey[i] = bkey_[i].ptr; f.bval[i] = bval_[i].ptr; }
        f.segA = segA_; f.segB = segB_; f.w = wv;
        f.ins_pos = ins_pos_.ptr; f.jobs = jobs_.ptr; f.njobs = counters_.ptr + 2; f.seg_d0 = seg_d0_.ptr; f.perm = perm_.ptr;
        f.wo = wvo; f.ins_posB = ins_posB_.ptr; f.jobsB = jobsB_.ptr; f.njobsB = counters_.ptr + 5;
        f.pend_prev = pend_prev; f.npend_prev = npend_prev;
        win_fused_small<<<1, kFusedThreads, 0, st_>>>(f); ++launch_counter();
        const int res = ((key_bits_ + 7) / 8) & 1;  // same parity rule as sort_pairs
        sa_key_ = akey_[res].ptr; sa_val_ = aval_[res].ptr;
        if (D_ == 1) { sb_key_ = bkey_[res].ptr; sb_val_ = bval_[res].ptr; }
        else { sb_key_ = sa_key_; sb_val_ = sa_val_; }
        log_start_ = (log_start_ + B) % W_;
        DPPR_CUDA(cudaGetLastError());
        record(2);
        batch_pending_ = true;
        return;
    }
    if (nA <= kCoopMaxEntries && coop_win_grid_ > 0 && tn_.window_path != 1) {
        // mid-size batch: the same stages inside one cooperative launch (window_coop.cuh)
        CoopArgs c{};
        c.log = log_.ptr; c.W = W_; c.log_start = log_start_; c.arriving = arriving; c.B = B;
        c.directed = D_ == 1; c.key_bits = key_bits_;
        for (int i = 0; i < 2; ++i) { c.akey[i] = akey_[i].ptr; c.aval[i] = aval_[i].ptr; c.bkey[i] = bkey_[i].ptr; c.bval[i] = bval_[i].ptr; }
        c.hist = sort_scratch_.ptr; c.tile_heads = tile_heads_.ptr;
        c.segA = segA_; c.segB = segB_; c.w = wv;
        c.ins_pos = ins_pos_.ptr; c.jobs = jobs_.ptr; c.njobs = counters_.ptr + 2; c.seg_d0 = seg_d0_.ptr; c.perm = perm_.ptr;
        c.wo = wvo; c.ins_posB = ins_posB_.ptr; c.jobsB = jobsB_.ptr; c.njobsB = counters_.ptr + 5;
        c.bar = counters_.ptr + 4;
        c.pend_prev = pend_prev; c.npend_prev = npend_prev;
        DPPR_CUDA(cudaMemsetAsync(counters_.ptr + 4, 0, sizeof(uint32_t), st_));
        const int tiles = div_up(nA, kSortTile);
        const int grid = std::max(1, std::min(coop_win_grid_, std::max(tiles, div_up(nA, kThreads))));
        void *params[] = {(void *)&c};
        DPPR_CUDA(cudaLaunchCooperativeKernel((void *)win_update_coop, dim3(grid), dim3(kThreads), params, 0, st_));
        ++launch_counter();
        const int res = ((key_bits_ + 7) / 8) & 1;
        sa_key_ = akey_[res].ptr; sa_val_ = aval_[res].ptr;
        if (D_ == 1) { sb_key_ = bkey_[res].ptr; sb_val_ = bval_[res].ptr; }
        else { sb_key_ = sa_key_; sb_val_ = sa_val_; }
        log_start_ = (log_start_ + B) % W_;
        record(2);
        batch_pending_ = true;
        return;
    }
    win_batch_entries<<<grid_for(B), kThreads, 0, st_>>>(log_.ptr, W_, log_start_, arriving, B, D_ == 1, V_,
                                                        akey_[0].ptr, aval_[0].ptr, bkey_[0].ptr, bval_[0].ptr, werr, perm_.ptr,
                                                        pend_prev, npend_prev, fr); ++launch_counter();
    log_start_ = (log_start_ + B) % W_;
    uint32_t *scan_scratch = sort_scratch_.ptr + sort_scratch_elems(Nb_);

    // group A: keyed by destination -> in-lists
    int res = sort_pairs(akey_[0].ptr, aval_[0].ptr, akey_[1].ptr, aval_[1].ptr, nA, key_bits_, sort_scratch_.ptr, st_);
    sa_key_ = akey_[res].ptr; sa_val_ = aval_[res].ptr;
    rle_heads<<<grid_for(nA), kThreads, 0, st_>>>(sa_key_, nA, flags_.ptr); ++launch_counter();
    exclusive_scan<uint32_t>(flags_.ptr, segA_.segof, nA, scan_scratch, nullptr, st_);
    rle_fill<<<grid_for(nA), kThreads, 0, st_>>>(sa_key_, sa_val_, nA, segA_); ++launch_counter();
    win_plan<<<grid_for(nA), kThreads, 0, st_>>>(segA_, wv, ins_pos_.ptr, jobs_.ptr, counters_.ptr + 2); ++launch_counter();
    win_relocate<<<std::min(grid_for(nA), 4 * sm_count_), kThreads, 0, st_>>>(jobs_.ptr, counters_.ptr + 2, pool_.ptr); ++launch_counter();
    win_insert<<<grid_for(nA), kThreads, 0, st_>>>(sa_key_, sa_val_, nA, segA_, ins_pos_.ptr, wv); ++launch_counter();

    // group B: keyed by source -> out-degrees + residual repair (undirected: same runs as group A)
    if (D_ == 1) {
        res = sort_pairs(bkey_[0].ptr, bval_[0].ptr, bkey_[1].ptr, bval_[1].ptr, nA, key_bits_, sort_scratch_.ptr, st_);
        sb_key_ = bkey_[res].ptr; sb_val_ = bval_[res].ptr;
        rle_heads<<<grid_for(nA), kThreads, 0, st_>>>(sb_key_, nA, flags_.ptr); ++launch_counter();
        exclusive_scan<uint32_t>(flags_.ptr, segB_.segof, nA, scan_scratch, nullptr, st_);
        rle_fill<<<grid_for(nA), kThreads, 0, st_>>>(sb_key_, sb_val_, nA, segB_); ++launch_counter();
    } else {
        sb_key_ = sa_key_; sb_val_ = sa_val_;
    }
    win_out_degrees<<<grid_for(nA), kThreads, 0, st_>>>(segB_, outdeg_.ptr, seg_d0_.ptr); ++launch_counter();
    if (outlists_) {  // out-lists: the same plan / relocate / insert on the source-sorted entries
        DPPR_CUDA(cudaMemsetAsync(counters_.ptr + 5, 0, sizeof(uint32_t), st_));
        win_plan<<<grid_for(nA), kThreads, 0, st_>>>(segB_, wvo, ins_posB_.ptr, jobsB_.ptr, counters_.ptr + 5); ++launch_counter();
        win_relocate<<<std::min(grid_for(nA), 4 * sm_count_), kThreads, 0, st_>>>(jobsB_.ptr, counters_.ptr + 5, pool_.ptr); ++launch_counter();
        win_insert<<<grid_for(nA), kThreads, 0, st_>>>(sb_key_, sb_val_, nA, segB_, ins_posB_.ptr, wvo); ++launch_counter();
    }
    DPPR_CUDA(cudaGetLastError());
    record(2);
    batch_pending_ = true;
}

// ---------------------------------------------------------------------------------------------
// one batch: residual repair + both push phases (the reference's timed region)
// ---------------------------------------------------------------------------------------------
void Engine::refresh(bool repair_only) {
    if (!batch_pending_) throw StateError("dppr_refresh without a preceding dppr_apply_batch");
    DPPR_CUDA(cudaSetDevice(dev_));
    const int64_t n = cur().entries;
    for (int k = 0; k < n_panels_; ++k) {  // (delta_ is one panel wide: repair_finalize leaves it cleared for the next)
        const size_t pb = (size_t)k * (size_t)panel_stride_;
        const int Sk = panel_sources(k);
        if (Pw_ == 1) {
            repair_accumulate<<<std::min(grid_for(n), 8 * sm_count_), kThreads, 0, st_>>>(sb_val_, n, segB_.segof, p_.ptr + pb, delta_.ptr); ++launch_counter();
        } else {
            dim3 g((unsigned)std::max(1, std::min(div_up(n, 32 * kWarps), 8 * sm_count_)), (unsigned)div_up(Sk, 32 * kRepairCols));
            repair_accumulate_rows<<<g, kThreads, 0, st_>>>(sb_val_, n, segB_.segof, p_.ptr + pb, Sr_, Sk, delta_.ptr); ++launch_counter();
        }
        repair_finalize<<<grid_for(n * Sk), kThreads, 0, st_>>>(segB_, seg_d0_.ptr, src_.ptr + (size_t)k * Pw_, Sk, Sr_, p_.ptr + pb, r_.ptr + pb,
                                                              delta_.ptr, cfg_.alpha); ++launch_counter();
    }
    DPPR_CUDA(cudaGetLastError());
    record(3);
    if (!repair_only) launch_push(false);
    else {
        DPPR_CUDA(cudaMemsetAsync(ctrl_.ptr, 0, kCtrlZeroBytes, st_));
        if (n_panels_ > 1) DPPR_CUDA(cudaMemsetAsync(ctrl_acc_.ptr, 0, sizeof(PushCtrl), st_));
    }
    record(4);
    finish_record();
    batch_pending_ = false;
}

void Engine::sync() {
    DPPR_CUDA(cudaSetDevice(dev_));
    DPPR_CUDA(cudaStreamSynchronize(st_));
    check_health();
}

void Engine::wait_event(void *cuda_event) {
    if (!cuda_event) throw InvalidArgument("null event");
    DPPR_CUDA(cudaSetDevice(dev_));
    DPPR_CUDA(cudaStreamWaitEvent(st_, (cudaEvent_t)cuda_event, 0));
}

// (does not fail on a flagged engine: the caller reads the flags here)
void Engine::get_stats(int64_t batch_index, dppr_batch_stats *out) {
    if (!out) throw InvalidArgument("null stats pointer");
    if (meta_.empty()) throw StateError("no batch has been processed yet");
    if (batch_index < 0) batch_index = (int64_t)meta_.size() - 1;
    if (batch_index >= (int64_t)meta_.size()) throw InvalidArgument("batch index out of range");
    if (batch_index == (int64_t)meta_.size() - 1 && batch_pending_) throw StateError("batch applied but not refreshed yet");
    DPPR_CUDA(cudaSetDevice(dev_));
    DPPR_CUDA(cudaStreamSynchronize(st_));
    BatchMeta &m = meta_[(size_t)batch_index];
    const BatchRecord *rec = record_slot((size_t)batch_index);
    if (cfg_.record_timing && !m.resolved) {
        EventSet &es = events_[batch_index % kEventSlots];
        if (es.batch == batch_index) resolve_events(es);
    }
    std::memset(out, 0, sizeof(*out));
    out->batch_index = batch_index;
    out->edges = m.edges;
    out->batch_entries = m.entries;  // N_b = 2*D*B
    out->touched_vertices = (D_ == 1) ? rec->nseg_out : rec->nseg_in;
    out->iterations = (int64_t)rec->iters;
    out->frontier_pops = (int64_t)rec->pops - (int64_t)rec->carried;  // carried items are not pushed
    out->traversed_edges = (int64_t)(rec->edges + rec->gath);
    out->scatter_edges = (int64_t)rec->edges;
    out->dense_slots = (int64_t)rec->walk_slots;
    out->dense_pairs = (int64_t)rec->walk_pairs;
    out->dense_units = (int64_t)rec->units;
    out->dense_pops = (int64_t)rec->dpops;
    out->hub_pops = (int64_t)rec->hubs;
    out->relocations = rec->njobs;
    out->pool_used = (int64_t)rec->pool_top;
    out->pool_leaked = (int64_t)rec->pool_leaked;
    out->error_flags = rec->errflags;
    out->dense_sweeps = (int32_t)rec->sweeps;
    out->ms_upload = m.ms[0];
    out->ms_window = m.ms[1];
    out->ms_repair = m.ms[2];
    out->ms_push = m.ms[3];
}

void Engine::get_vector(int which, int32_t s, double *out) {
    if (!out) throw InvalidArgument("null output pointer");
    if (s < 0 || s >= S_) throw InvalidArgument("source index out of range");
    if (!solved_) throw StateError("no estimates before dppr_solve_initial");
    sync();
    const double *src = (which == 0 ? p_.ptr : r_.ptr) + elem_base(s);  // vertex-major inside the source's panel: element v at + v * Sr_
    if (!perm_.ptr && Sr_ == 1) {
        DPPR_CUDA(cudaMemcpy(out, src, sizeof(double) * (size_t)V_, cudaMemcpyDeviceToHost));
        return;
    }
    DevBuf<double> tmp;
    tmp.alloc((size_t)V_);
    gather_by_perm<double><<<grid_for(V_), kThreads, 0, st_>>>(src, Sr_, perm_.ptr, tmp.ptr, V_); ++launch_counter();
    DPPR_CUDA(cudaStreamSynchronize(st_));
    DPPR_CUDA(cudaMemcpy(out, tmp.ptr, sizeof(double) * (size_t)V_, cudaMemcpyDeviceToHost));
}

void Engine::copy_estimates_device(int32_t s, void *dptr) {
    if (!dptr) throw InvalidArgument("null device pointer");
    if (s < 0 || s >= S_) throw InvalidArgument("source index out of range");
    DPPR_CUDA(cudaSetDevice(dev_));
    gather_by_perm<double><<<grid_for(V_), kThreads, 0, st_>>>(p_.ptr + elem_base(s), Sr_, perm_.ptr, (double *)dptr, V_); ++launch_counter();
    DPPR_CUDA(cudaStreamSynchronize(st_));
}

void Engine::set_state(int32_t s, const double *p, const double *r) {
    if (s < 0 || s >= S_) throw InvalidArgument("source index out of range");
    sync();
    DevBuf<double> tmp;
    tmp.alloc((size_t)V_);
    for (int which = 0; which < 2; ++which) {
        const double *h = which == 0 ? p : r;
        if (!h) continue;
        double *dst = (which == 0 ? p_.ptr : r_.ptr) + elem_base(s);
        // (on the engine's stream: a pageable cudaMemcpy on the legacy stream may return before its DMA has finished, and the
        // engine's non-blocking stream does not order against it)
        DPPR_CUDA(cudaMemcpyAsync(tmp.ptr, h, sizeof(double) * (size_t)V_, cudaMemcpyHostToDevice, st_));
        scatter_by_perm<double><<<grid_for(V_), kThreads, 0, st_>>>(tmp.ptr, perm_.ptr, dst, Sr_, V_); ++launch_counter();
        DPPR_CUDA(cudaStreamSynchronize(st_));
    }
    solved_ = true;
    if (meta_.empty()) meta_.emplace_back();
}

int Engine::get_iterlog(uint32_t *out, int cap) {
    if (!iterlog_.ptr || meta_.empty()) return 0;
    sync();
    const BatchRecord *rec = record_slot(meta_.size() - 1);
    int n = (int)std::min<unsigned long long>(rec->iters, (unsigned long long)std::min(cap, kIterLogCap));
    if (n > 0) DPPR_CUDA(cudaMemcpy(out, iterlog_.ptr, sizeof(uint4) * (size_t)n, cudaMemcpyDeviceToHost));
    return n;
}

int Engine::get_ctalog(unsigned long long *out, int cap_rows) {
    if (!ctalog_.ptr) return 0;
    sync();
    int rows = std::min(cap_rows, coop_grid_[cfg_.variant]);
    DPPR_CUDA(cudaMemcpy(out, ctalog_.ptr, sizeof(unsigned long long) * 8 * (size_t)rows, cudaMemcpyDeviceToHost));
    return rows;
}

// canonical CSR: rows ascending, duplicates kept.  Device sort, test / validation path.
void Engine::export_sorted_device(SortedCsr &o, bool out_lists) {
    if (!window_ready_) throw StateError("dppr_export_window_csr before dppr_init_window");
    if (out_lists && !outlists_ && D_ == 1) throw StateError("this engine does not maintain out-lists (dense iterations are off)");
    const uint4 *vm = (out_lists && outlists_) ? vmeta_out_.ptr : vmeta_.ptr;
    sync();
    o.len.alloc((size_t)V_); o.rowptr.alloc((size_t)V_ + 1);
    for (int i = 0; i < 2; ++i) { o.key[i].alloc((size_t)Ew_); o.val[i].alloc((size_t)Ew_); }
    o.scratch.alloc(std::max(sort_scratch_elems(Ew_), scan_scratch_elems(V_)));
    o.total.alloc(1);
    win_export_len<<<grid_for(V_), kThreads, 0, st_>>>(vm, o.len.ptr, V_, perm_.ptr); ++launch_counter();
    exclusive_scan<uint32_t>(o.len.ptr, o.rowptr.ptr, V_, o.scratch.ptr, o.total.ptr, st_);
    DPPR_CUDA(cudaMemcpyAsync(o.rowptr.ptr + V_, o.total.ptr, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st_));
    uint32_t htotal = 0;
    DPPR_CUDA(cudaMemcpyAsync(&htotal, o.total.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st_));
    DPPR_CUDA(cudaStreamSynchronize(st_));
    if ((int64_t)htotal != Ew_)
        throw StateError("window graph holds " + std::to_string(htotal) + " entries, expected " + std::to_string(Ew_));
    win_export_entries<<<grid_for((int64_t)V_ * 32), kThreads, 0, st_>>>(vm, pool_.ptr, o.rowptr.ptr, o.key[0].ptr,
                                                                      o.val[0].ptr, V_, perm_.ptr, inv_.ptr); ++launch_counter();
    // sort by (dst, src): LSD over the pair = stable sort by src, then stable sort by dst
    int res = sort_pairs(o.val[0].ptr, o.key[0].ptr, o.val[1].ptr, o.key[1].ptr, Ew_, key_bits_, o.scratch.ptr, st_);
    uint32_t *k0 = o.key[res].ptr, *v0 = o.val[res].ptr, *k1 = o.key[1 - res].ptr, *v1 = o.val[1 - res].ptr;
    res = sort_pairs(k0, v0, k1, v1, Ew_, key_bits_, o.scratch.ptr, st_);
    o.rows = res ? k1 : k0;
    o.cols = res ? v1 : v0;
    DPPR_CUDA(cudaGetLastError());
    DPPR_CUDA(cudaStreamSynchronize(st_));
}

void Engine::export_csr(int32_t *in_row_ptr, int32_t *in_col_ind, int32_t *out_deg, bool out_lists) {
    SortedCsr o;
    export_sorted_device(o, out_lists);
    if (in_row_ptr) DPPR_CUDA(cudaMemcpy(in_row_ptr, o.rowptr.ptr, sizeof(int32_t) * ((size_t)V_ + 1), cudaMemcpyDeviceToHost));
    if (in_col_ind && Ew_ > 0) DPPR_CUDA(cudaMemcpy(in_col_ind, o.cols, sizeof(int32_t) * (size_t)Ew_, cudaMemcpyDeviceToHost));
    if (out_deg) {
        DevBuf<int32_t> od;
        od.alloc((size_t)V_);
        gather_by_perm<int32_t><<<grid_for(V_), kThreads, 0, st_>>>(outdeg_.ptr, 1, perm_.ptr, od.ptr, V_); ++launch_counter();
        DPPR_CUDA(cudaStreamSynchronize(st_));
        DPPR_CUDA(cudaMemcpy(out_deg, od.ptr, sizeof(int32_t) * (size_t)V_, cudaMemcpyDeviceToHost));
    }
}

// the reference's ValidateGraph (gpu/PPRRevPushGPU.cuh:45-90) without the host: expected entries from the window's own
// edges, same stable sorts, compared on the device
void Engine::check_window_device(const int32_t *dpairs, int64_t n, int64_t *mismatches) {
    if (!dpairs || !mismatches) throw InvalidArgument("null argument");
    if (n != W_) throw InvalidArgument("dppr_check_window_device needs exactly window_edges pairs");
    SortedCsr mine;
    export_sorted_device(mine, false);
    DevBuf<uint32_t> key[2], val[2], od_exp, scratch;
    DevBuf<int32_t> od_mine;
    DevBuf<unsigned long long> bad;
    for (int i = 0; i < 2; ++i) { key[i].alloc((size_t)Ew_); val[i].alloc((size_t)Ew_); }
    od_exp.alloc((size_t)V_); od_mine.alloc((size_t)V_);
    scratch.alloc(sort_scratch_elems(Ew_));
    bad.alloc(1);
    DPPR_CUDA(cudaMemsetAsync(od_exp.ptr, 0, od_exp.bytes(), st_));
    DPPR_CUDA(cudaMemsetAsync(bad.ptr, 0, sizeof(unsigned long long), st_));
    val_window_entries<<<grid_for(W_), kThreads, 0, st_>>>((const int2 *)dpairs, W_, D_ == 1, V_, key[0].ptr, val[0].ptr, od_exp.ptr, bad.ptr); ++launch_counter();
    int res = sort_pairs(val[0].ptr, key[0].ptr, val[1].ptr, key[1].ptr, Ew_, key_bits_, scratch.ptr, st_);
    uint32_t *k0 = key[res].ptr, *v0 = val[res].ptr, *k1 = key[1 - res].ptr, *v1 = val[1 - res].ptr;
    res = sort_pairs(k0, v0, k1, v1, Ew_, key_bits_, scratch.ptr, st_);
    const uint32_t *rows = res ? k1 : k0, *cols = res ? v1 : v0;
    val_count_diff<<<grid_for(Ew_), kThreads, 0, st_>>>(rows, mine.rows, cols, mine.cols, Ew_, bad.ptr); ++launch_counter();
    gather_by_perm<int32_t><<<grid_for(V_), kThreads, 0, st_>>>(outdeg_.ptr, 1, perm_.ptr, od_mine.ptr, V_); ++launch_counter();
    val_count_diff<<<grid_for(V_), kThreads, 0, st_>>>(od_exp.ptr, (const uint32_t *)od_mine.ptr, nullptr, nullptr, V_, bad.ptr); ++launch_counter();
    DPPR_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    DPPR_CUDA(cudaMemcpyAsync(&h, bad.ptr, sizeof(h), cudaMemcpyDeviceToHost, st_));
    DPPR_CUDA(cudaStreamSynchronize(st_));
    *mismatches = (int64_t)h;
}

void Engine::validate(int32_t s, double *max_abs_residual, double *max_invariant_defect) {
    if (s < 0 || s >= S_) throw InvalidArgument("source index out of range");
    if (!solved_) throw StateError("no estimates before dppr_solve_initial");
    sync();
    DevBuf<unsigned long long> out;
    out.alloc(2);
    DPPR_CUDA(cudaMemsetAsync(out.ptr, 0, out.bytes(), st_));
    const double *p = p_.ptr + elem_base(s), *r = r_.ptr + elem_base(s);
    val_residual_max<<<grid_for(V_), kThreads, 0, st_>>>(r, Sr_, V_, out.ptr); ++launch_counter();
    DevBuf<double> acc;
    if (max_invariant_defect) {
        acc.alloc((size_t)V_);
        DPPR_CUDA(cudaMemsetAsync(acc.ptr, 0, acc.bytes(), st_));
        val_out_sums<<<grid_for((int64_t)V_ * 32), kThreads, 0, st_>>>(vmeta_.ptr, pool_.ptr, p, Sr_, V_, acc.ptr); ++launch_counter();
        int32_t hsrc = 0;
        DPPR_CUDA(cudaMemcpyAsync(&hsrc, src_.ptr + s, sizeof(int32_t), cudaMemcpyDeviceToHost, st_));
        DPPR_CUDA(cudaStreamSynchronize(st_));
        val_invariant<<<grid_for(V_), kThreads, 0, st_>>>(p, r, Sr_, outdeg_.ptr, acc.ptr, V_, hsrc, cfg_.alpha, out.ptr + 1); ++launch_counter();
    }
    DPPR_CUDA(cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    DPPR_CUDA(cudaMemcpyAsync(h, out.ptr, sizeof(h), cudaMemcpyDeviceToHost, st_));
    DPPR_CUDA(cudaStreamSynchronize(st_));
    double d[2];
    std::memcpy(d, h, sizeof(d));
    if (max_abs_residual) *max_abs_residual = d[0];
    if (max_invariant_defect) *max_invariant_defect = d[1];
}

// the k largest estimates of sources [first, first + n): csrc/topk.cuh
void Engine::topk(int32_t first, int32_t n, int32_t k, int32_t *ids, double *values) {
    if (!ids || !values) throw InvalidArgument("null output pointer");
    if (first < 0 || n < 1 || first + n > S_) throw InvalidArgument("source range out of bounds");
    if (k < 1 || k > kTopKMax) throw InvalidArgument("k must be in [1, 128]");
    if (!solved_) throw StateError("no estimates before dppr_solve_initial");
    DPPR_CUDA(cudaSetDevice(dev_));
    const size_t outn = (size_t)n * k;
    if (topk_out_ids_.count < outn) {
        topk_out_ids_.alloc(outn); topk_out_vals_.alloc(outn);
        topk_host_ids_.alloc(outn); topk_host_vals_.alloc(outn);
    }
    if (!topk_prevk_.ptr) {  // per-engine state of the fast path (topk.cuh)
        topk_prev_.alloc((size_t)S_ * kTopKMax); topk_prevk_.alloc(S_); topk_bound_.alloc(S_); topk_count_.alloc(S_);
        topk_over_.alloc(S_); topk_host_over_.alloc(S_);
        topk_ckey_.alloc((size_t)S_ * kTopCand); topk_cid_.alloc((size_t)S_ * kTopCand);
        DPPR_CUDA(cudaMemsetAsync(topk_prevk_.ptr, 0, topk_prevk_.bytes(), st_));
        DPPR_CUDA(cudaMemsetAsync(topk_count_.ptr, 0, topk_count_.bytes(), st_));
    }
    // the request's sources, panel by panel: [lf, lf + ln) inside panel kp, at offset `off` of the request-relative arrays
    struct Part { int kp, lf, ln, off; };
    std::vector<Part> parts;
    for (int kp = first / Pw_; kp <= (first + n - 1) / Pw_; ++kp) {
        const int g0 = std::max(first, kp * Pw_), g1 = std::min(first + n, kp * Pw_ + panel_sources(kp));
        if (g1 > g0) parts.push_back({kp, g0 - kp * Pw_, g1 - g0, g0 - first});
    }
    const int64_t rows_per_warp = Sr_ < 32 ? 32 / Sr_ : 1;
    const int fgrid = (int)std::max<int64_t>(1, std::min<int64_t>(div_up(div_up((int64_t)V_, rows_per_warp), (int64_t)(kThreads / 32)), (int64_t)sm_count_ * 32));
    for (const Part &q : parts)
        topk_threshold<<<(unsigned)q.ln, kThreads, 0, st_>>>(p_.ptr + (size_t)q.kp * panel_stride_, Sr_, V_, q.lf, k,
                                                            topk_prev_.ptr + (size_t)q.kp * Pw_ * kTopKMax, topk_prevk_.ptr + (size_t)q.kp * Pw_,
                                                            topk_bound_.ptr + q.off), ++launch_counter();
    bool open_sources = true;
    for (int round = 0; round < 3 && open_sources; ++round) {
        for (const Part &q : parts) {
            const int gfirst = q.kp * Pw_ + q.lf;
            topk_filter<<<fgrid, kThreads, 0, st_>>>(p_.ptr + (size_t)q.kp * panel_stride_, Sr_, V_, q.lf, q.ln, inv_.ptr, topk_bound_.ptr + q.off,
                                                     topk_ckey_.ptr + (size_t)q.off * kTopCand, topk_cid_.ptr + (size_t)q.off * kTopCand,
                                                     topk_count_.ptr + q.off, round ? topk_over_.ptr + q.off : nullptr); ++launch_counter();
            topk_merge<<<(unsigned)q.ln, kThreads, 0, st_>>>(topk_ckey_.ptr + (size_t)q.off * kTopCand, topk_cid_.ptr + (size_t)q.off * kTopCand, kTopCand, 0,
                                                            topk_count_.ptr + q.off, k, topk_out_ids_.ptr + (size_t)q.off * k,
                                                            topk_out_vals_.ptr + (size_t)q.off * k, perm_.ptr,
                                                            topk_prev_.ptr + (size_t)gfirst * kTopKMax, topk_prevk_.ptr + gfirst, topk_over_.ptr + q.off,
                                                            round > 0, topk_bound_.ptr + q.off); ++launch_counter();
        }
        DPPR_CUDA(cudaGetLastError());
        DPPR_CUDA(cudaMemcpyAsync(topk_host_over_.ptr, topk_over_.ptr, sizeof(int) * n, cudaMemcpyDeviceToHost, st_));
        if (round == 0) {  // (the common case ends here: one synchronisation for flags and results)
            DPPR_CUDA(cudaMemcpyAsync(topk_host_ids_.ptr, topk_out_ids_.ptr, sizeof(int32_t) * outn, cudaMemcpyDeviceToHost, st_));
            DPPR_CUDA(cudaMemcpyAsync(topk_host_vals_.ptr, topk_out_vals_.ptr, sizeof(double) * outn, cudaMemcpyDeviceToHost, st_));
        }
        DPPR_CUDA(cudaStreamSynchronize(st_));
        open_sources = false;
        for (int32_t i = 0; i < n; ++i) open_sources = open_sources || topk_host_over_.ptr[i] != 0;
        if (open_sources && debug_) {
            int cnt = 0, mx = 0;
            for (int32_t i = 0; i < n; ++i) { cnt += topk_host_over_.ptr[i] != 0; mx = std::max(mx, topk_host_over_.ptr[i]); }
            std::fprintf(stderr, "[dppr] top-k round %d: %d of %d sources over the candidate capacity (largest list %d)\n", round, cnt, n, mx);
        }
        if (round == 0 && !open_sources) { check_health(); std::memcpy(ids, topk_host_ids_.ptr, sizeof(int32_t) * outn); std::memcpy(values, topk_host_vals_.ptr, sizeof(double) * outn); return; }
    }
    check_health();
    // sources three rounds did not settle (a flat vector: the bound cannot rise past a tie): the exact scan
    for (int32_t i = 0; i < n && open_sources; ++i) {
        if (!topk_host_over_.ptr[i]) continue;
        const int gs = first + i;  // (global source index)
        const int slices = div_up(V_, kTopSlice);
        const size_t need = (size_t)slices * k;
        if (topk_key_.count < need) { topk_key_.alloc(need); topk_id_.alloc(need); }
        topk_partial<<<dim3((unsigned)slices, 1u), kThreads, 0, st_>>>(p_.ptr + (size_t)(gs / Pw_) * panel_stride_, Sr_, V_, gs % Pw_, inv_.ptr, k, topk_key_.ptr,
                                                                       topk_id_.ptr); ++launch_counter();
        topk_merge<<<1, kThreads, 0, st_>>>(topk_key_.ptr, topk_id_.ptr, (int64_t)need, (int64_t)need, nullptr, k, topk_out_ids_.ptr + (size_t)i * k,
                                           topk_out_vals_.ptr + (size_t)i * k, perm_.ptr, topk_prev_.ptr + (size_t)gs * kTopKMax,
                                           topk_prevk_.ptr + gs, nullptr, 0, nullptr); ++launch_counter();
    }
    DPPR_CUDA(cudaGetLastError());
    DPPR_CUDA(cudaMemcpyAsync(topk_host_ids_.ptr, topk_out_ids_.ptr, sizeof(int32_t) * outn, cudaMemcpyDeviceToHost, st_));
    DPPR_CUDA(cudaMemcpyAsync(topk_host_vals_.ptr, topk_out_vals_.ptr, sizeof(double) * outn, cudaMemcpyDeviceToHost, st_));
    DPPR_CUDA(cudaStreamSynchronize(st_));
    std::memcpy(ids, topk_host_ids_.ptr, sizeof(int32_t) * outn);
    std::memcpy(values, topk_host_vals_.ptr, sizeof(double) * outn);
}

}  // namespace dppr
