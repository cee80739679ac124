// engine.cu -- see engine.cuh.  Reference call stack being replaced: SURVEY.md 3.1 / 3.2.
#include "engine.cuh"
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstring>

namespace dppr {

namespace {

__global__ void gather_record(BatchRecord *rec, const PushCtrl *ctrl, const uint32_t *counters,
                              const unsigned long long *pool_top) {
    rec->ctrl = *ctrl;
    rec->nseg_in = counters[0];
    rec->nseg_out = counters[1];
    rec->njobs = counters[2] + counters[5];
    rec->pad = 0;
    rec->pool_top = *pool_top;
}

__global__ void fold_window_errors(PushCtrl *ctrl, int *win_err) {
    if (*win_err) atomicOr(&ctrl->errflags, *win_err);
}

int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

template <int VAR, bool DENSE = false>
void *persistent_kernel() { return (void *)push_persistent<VAR, DENSE>; }

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
    if (cfg.engine_mode < DPPR_ENGINE_AUTO || cfg.engine_mode > DPPR_ENGINE_LEVELSYNC)
        throw InvalidArgument("engine_mode must be one of DPPR_ENGINE_{AUTO,STEPWISE,ASYNC,LEVELSYNC}");
    if (cfg.engine_mode == DPPR_ENGINE_ASYNC && cfg.variant != DPPR_OPTIMIZED)
        throw InvalidArgument("DPPR_ENGINE_ASYNC implements variant 0 (optimized) only");
    mode_ = cfg.engine_mode;
    if (mode_ == DPPR_ENGINE_AUTO) mode_ = DPPR_ENGINE_LEVELSYNC;  // measured fastest for every variant so far (profiles/README.md)
    if (const char *force = std::getenv("DPPR_FORCE_ENGINE")) {  // tuning / A-B runs without touching the caller
        const int f = std::atoi(force);
        if (f >= DPPR_ENGINE_STEPWISE && f <= DPPR_ENGINE_LEVELSYNC && !(f == DPPR_ENGINE_ASYNC && cfg.variant != 0)) mode_ = f;
    }
    if (mode_ == DPPR_ENGINE_ASYNC && cfg.n_sources > kMaxAsyncSources)
        throw InvalidArgument("DPPR_ENGINE_ASYNC supports at most 4096 sources per engine");
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
    Vp_ = ((int64_t)V_ + 31) / 32 * 32;
    sources_.assign(cfg.sources, cfg.sources + S_);
    cfg_.sources = sources_.data();
    for (int32_t s : sources_)
        if (s < 0 || s >= V_) throw InvalidArgument("source vertex id out of range");
    if ((uint64_t)V_ >= (1ull << 31)) throw InvalidArgument("vertex ids must fit in 31 bits");
    if (Ew_ >= (int64_t)0xffffffffll) throw InvalidArgument("window too large for 32-bit CSR offsets");
    key_bits_ = bits_for((uint64_t)(V_ - 1));
    relabel_ = env_int("DPPR_RELABEL", 1) != 0;
    {
        // dense iterations (pull.cuh): variant 0 of the level-synchronous engine; an iteration runs as a gather sweep
        // once it is expected to traverse at least (E_w + 2 V) x sources / DPPR_DENSE_DIV in-edges.  0 disables.
        const char *dd = std::getenv("DPPR_DENSE_DIV");
        dense_div_ = dd ? std::atof(dd) : 4.0;
        // Small windows stay with the scatter-only kernel: their iterations are bound by the chain of dependent round
        // trips, which a sweep does not shorten (BASELINE configs[1]: 43 vs 18 us), and the kernel that can switch
        // carries more loop state, which costs its scatter iterations +9..+24 % when they are latency-bound (LJ/4,
        // youtube) and nothing when they are bandwidth-bound (Twitter-shaped).
        const char *me = std::getenv("DPPR_DENSE_MIN_EDGES");
        const double min_edges = me ? std::atof(me) : 2.0e7;
        dense_ = dense_div_ > 0.0 && cfg.variant == DPPR_OPTIMIZED && mode_ == DPPR_ENGINE_LEVELSYNC &&
                 (double)cfg.window_edges * (cfg.directed ? 1 : 2) * cfg.n_sources >= min_edges;
        outlists_ = dense_ && D_ == 1;
        // several sources: rows of x hold 4-source chunks, G = 2^gshift adjacent lanes take G chunks of a vertex (pull.cuh)
        pull_gshift_ = 0;
        if (S_ > 1) {
            const int chunks = (S_ + 3) / 4;
            const int gmax = std::min(std::max(env_int("DPPR_PULL_GROUP", 8), 1), 8);
            while ((2 << pull_gshift_) <= std::min(chunks, gmax)) ++pull_gshift_;
        }
        Sp_ = S_ == 1 ? 1 : (S_ + (4 << pull_gshift_) - 1) / (4 << pull_gshift_) * (4 << pull_gshift_);
        pull_warp_min_ = std::max(1, env_int("DPPR_PULL_WARP_MIN", 32));
        pull_cta_min_ = std::max(pull_warp_min_, env_int("DPPR_PULL_CTA_MIN", 1024));
        pull_big_min_ = std::max(pull_cta_min_, env_int("DPPR_PULL_BIG_MIN", 65536));
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
    void *kern[4] = {dense_ ? persistent_kernel<0, true>() : persistent_kernel<0>(), persistent_kernel<1>(),
                     persistent_kernel<2>(), persistent_kernel<3>()};
    const int want_per_sm = env_int("DPPR_CTAS_PER_SM", 4);
    for (int v = 0; v < 4; ++v) {
        int per_sm = 0;
        DPPR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern[v], kThreads, 0));
        if (per_sm < 1) throw CudaFailure("push kernel does not fit on an SM");
        coop_grid_[v] = std::min(per_sm, std::max(want_per_sm, 1)) * sm_count_;
    }
    {
        int per_sm = 0;
        DPPR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (void *)push_async, kThreads, 0));
        if (per_sm < 1) throw CudaFailure("async push kernel does not fit on an SM");
        async_grid_ = std::min(per_sm, std::max(env_int("DPPR_ASYNC_CTAS_PER_SM", 4), 1)) * sm_count_;
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
    pool_top_.alloc(1);
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
            x_[i].alloc((size_t)V_ * Sp_);
            DPPR_CUDA(cudaMemsetAsync(x_[i].ptr, 0, x_[i].bytes(), st_));  // the padding columns stay zero for good
        }
        tile_list_.alloc((size_t)div_up(V_, kThreads >> pull_gshift_) * (Sp_ == 1 ? 1 : (Sp_ / 4) >> pull_gshift_));
        bigcap_ = (uint32_t)std::min<int64_t>((Ew_ / pull_big_min_ + 64) * (Sp_ == 1 ? 1 : Sp_ / 4), 1 << 24);
        big_.alloc(bigcap_);
        bigacc_.alloc((size_t)bigcap_ * 4 << pull_gshift_);
        DPPR_CUDA(cudaMemsetAsync(bigacc_.ptr, 0, bigacc_.bytes(), st_));
    }
    seg_d0_.alloc((size_t)Nb_);
    delta_.alloc((size_t)Nb_ * S_);
    DPPR_CUDA(cudaMemsetAsync(delta_.ptr, 0, delta_.bytes(), st_));
    DPPR_CUDA(cudaMemsetAsync(counters_.ptr, 0, counters_.bytes(), st_));
    // state
    p_.alloc((size_t)Vp_ * S_);
    r_.alloc((size_t)Vp_ * S_);
    if (cfg_.variant >= DPPR_EAGER) status_.alloc((size_t)Vp_ * S_);
    src_.alloc((size_t)S_);
    DPPR_CUDA(cudaMemcpyAsync(src_.ptr, sources_.data(), sizeof(int32_t) * S_, cudaMemcpyHostToDevice, st_));
    // push queues: a frontier holds each (source, vertex) at most once
    int64_t qc = cfg_.frontier_capacity > 0 ? cfg_.frontier_capacity : std::min<int64_t>((int64_t)V_ * S_, (int64_t)1 << 29);
    qc = std::max<int64_t>(qc, 1024);
    if (qc > 0xfffffff0ll) qc = 0xfffffff0ll;
    qcap_ = (uint32_t)qc;
    hcap_ = (uint32_t)std::min<int64_t>(qc, std::max<int64_t>(1024, (Ew_ / cfg_.hub_degree + 1) * S_));
    if (use_async()) {
        // ticket rings (one per phase): at most one outstanding entry per (source, vertex) plus hub chunks
        // ... plus the tickets idle warps hold ahead of `tail`: two live tickets must never share a slot
        const unsigned long long worst = 2ull * (unsigned long long)V_ * S_ + (unsigned long long)Ew_ / kHubChunk * S_ +
                                         64ull * (unsigned long long)async_grid_ * kWarps;
        unsigned long long want = cfg_.frontier_capacity > 0 ? (unsigned long long)cfg_.frontier_capacity
                                                             : std::min<unsigned long long>(worst, 1ull << 30);
        ring_cap_ = 1ull << 16;
        while (ring_cap_ < want) ring_cap_ <<= 1;
        guard_slots_ = ring_cap_ < worst ? 1 : 0;
        for (int i = 0; i < 2; ++i) {
            q_[i].alloc(ring_cap_);
            qr_[i].alloc(ring_cap_);
            DPPR_CUDA(cudaMemsetAsync(q_[i].ptr, 0xff, q_[i].bytes(), st_));  // every slot EMPTY
        }
        async_ctr_.alloc(10 * 16);
        DPPR_CUDA(cudaMemsetAsync(async_ctr_.ptr, 0, async_ctr_.bytes(), st_));
    } else {
        for (int i = 0; i < 2; ++i) {
            q_[i].alloc(qcap_);
            if (cfg_.variant != DPPR_OPTIMIZED) qr_[i].alloc(qcap_);
            hub_[i].alloc(hcap_);
        }
    }
    ctrl_.alloc(1);
    DPPR_CUDA(cudaMemsetAsync(ctrl_.ptr, 0, sizeof(PushCtrl), st_));
    dev_record_.alloc(1);
    if (env_int("DPPR_ITERLOG", 0)) {
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
    for (auto &m : meta_)
        for (auto &ev : m.ev)
            if (ev) cudaEventDestroy(ev);
    for (auto &ev : hstage_free_)
        if (ev) cudaEventDestroy(ev);
    if (st_) cudaStreamDestroy(st_);
}

void Engine::record(int which) {
    if (!cfg_.record_timing) return;
    BatchMeta &m = cur();
    if (!m.ev[which]) DPPR_CUDA(cudaEventCreate(&m.ev[which]));
    DPPR_CUDA(cudaEventRecord(m.ev[which], st_));
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
    gather_record<<<1, 1, 0, st_>>>(dev_record_.ptr, ctrl_.ptr, counters_.ptr, pool_top_.ptr); ++launch_counter();
    DPPR_CUDA(cudaGetLastError());
    DPPR_CUDA(cudaMemcpyAsync(record_slot(meta_.size() - 1), dev_record_.ptr, sizeof(BatchRecord),
                              cudaMemcpyDeviceToHost, st_));
}

// ---------------------------------------------------------------------------------------------
// initial window
// ---------------------------------------------------------------------------------------------
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
    int *werr = (int *)(counters_.ptr + 3);

    if (relabel_) {
        // internal order = descending out-degree of the initial window (window.cuh, "internal vertex order")
        DevBuf<uint32_t> deg, rk[2], rv[2], rscratch;
        deg.alloc((size_t)V_);
        for (int i = 0; i < 2; ++i) { rk[i].alloc((size_t)V_); rv[i].alloc((size_t)V_); }
        rscratch.alloc(sort_scratch_elems(V_));
        perm_.alloc((size_t)V_); inv_.alloc((size_t)V_);
        DPPR_CUDA(cudaMemsetAsync(deg.ptr, 0, deg.bytes(), st_));
        // key: out-degree (how often the scatter form hits r[w]); with dense iterations on a directed graph out- plus
        // in-degree (the gather form reads x[u] once per IN-edge of u): DPPR_RELABEL_BOTH=1; no difference measured on R-MAT
        const int both = env_int("DPPR_RELABEL_BOTH", 0);
        relabel_degrees<<<grid_for(W_), kThreads, 0, st_>>>(log_.ptr, W_, D_ == 1, V_, deg.ptr, werr, both); ++launch_counter();
        const int dbits = bits_for((uint64_t)2 * Ew_);
        const uint32_t degmax = (uint32_t)((1ull << dbits) - 1);
        relabel_keys<<<grid_for(V_), kThreads, 0, st_>>>(deg.ptr, degmax, rk[0].ptr, rv[0].ptr, V_); ++launch_counter();
        const int rr = sort_pairs(rk[0].ptr, rv[0].ptr, rk[1].ptr, rv[1].ptr, V_, dbits, rscratch.ptr, st_);
        const uint32_t P = (uint32_t)std::max(1, std::min(env_int("DPPR_RELABEL_BLOCKS", 1024), V_));
        relabel_assign<<<grid_for(V_), kThreads, 0, st_>>>(rv[rr].ptr, perm_.ptr, inv_.ptr, V_, P); ++launch_counter();
        relabel_log<<<grid_for(W_), kThreads, 0, st_>>>(log_.ptr, W_, perm_.ptr, V_); ++launch_counter();
        // the sources, in internal ids
        std::vector<uint32_t> hp((size_t)S_);
        DPPR_CUDA(cudaStreamSynchronize(st_));
        for (int s = 0; s < S_; ++s)
            DPPR_CUDA(cudaMemcpy(&hp[s], perm_.ptr + sources_[s], sizeof(uint32_t), cudaMemcpyDeviceToHost));
        DPPR_CUDA(cudaMemcpy(src_.ptr, hp.data(), sizeof(uint32_t) * (size_t)S_, cudaMemcpyHostToDevice));
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
    window_ready_ = true;
    solved_ = false;
    batch_pending_ = false;
}

// ---------------------------------------------------------------------------------------------
// push launch
// ---------------------------------------------------------------------------------------------
void Engine::launch_push(bool init_mode) {
    DPPR_CUDA(cudaMemsetAsync(ctrl_.ptr, 0, kCtrlZeroBytes, st_));
    PushArgs a{};
    a.vmeta = vmeta_.ptr; a.pool = pool_.ptr; a.outdeg = outdeg_.ptr;
    a.p = p_.ptr; a.r = r_.ptr; a.status = status_.ptr;
    a.Vp = Vp_; a.S = S_; a.src = src_.ptr;
    for (int i = 0; i < 2; ++i) { a.q[i] = q_[i].ptr; a.qr[i] = qr_[i].ptr; a.hub[i] = hub_[i].ptr; }
    a.qcap = qcap_; a.hcap = hcap_;
    a.cand = segB_.vertex; a.ncand = segB_.count;
    a.ctrl = ctrl_.ptr;
    a.eps = cfg_.epsilon; a.alpha = cfg_.alpha;
    a.hub_degree = cfg_.hub_degree;
    a.init_mode = init_mode ? 1 : 0;
    a.max_iters = env_int("DPPR_MAX_ITERS", 400000);
    {
        const char *g = std::getenv("DPPR_CARRY_GAMMA"), *sc = std::getenv("DPPR_CARRY_SCALE");
        // off by default: on the L2-resident BASELINE configs the extra (latency-bound) iterations cost more than
        // the saved traversals (profiles/README.md); DPPR_CARRY_GAMMA=0.7 enables the threshold schedule
        a.carry_gamma = g ? std::atof(g) : 1.0;
        a.carry_scale = sc ? std::atof(sc) : 0.01;
    }
    a.tile_cap = std::min(std::max(env_int("DPPR_TILE_CAP", 128), 8), kTileMax);
    a.V = V_;
    a.avg_indeg = (float)((double)Ew_ / (double)V_);
    a.vmeta_out = outlists_ ? vmeta_out_.ptr : vmeta_.ptr;
    a.x[0] = x_[0].ptr; a.x[1] = x_[1].ptr;
    a.Sp = Sp_; a.pull_gshift = pull_gshift_;
    // cost model: a sweep reads every out-list entry and every vertex row once, whatever the frontier; a scatter
    // iteration pays one random atomic per traversed in-edge.  Measured ratio ~ DPPR_DENSE_DIV (3): Twitter-shaped
    // 3.3 ms per sweep vs 17 edges/ns scattered; Orkut/4 97 us vs 40 edges/ns.
    a.dense_enter_edges = ~0ull;
    if (dense_) a.dense_enter_edges = (unsigned long long)std::max(1.0, ((double)Ew_ + 2.0 * (double)V_) * (double)S_ / dense_div_);
    a.dense_exit_edges = a.dense_enter_edges / 2;
    a.pull_warp_min = pull_warp_min_; a.pull_cta_min = pull_cta_min_; a.pull_big_min = pull_big_min_;
    a.big = big_.ptr; a.bigcap = bigcap_; a.bigacc = bigacc_.ptr; a.tile_list = tile_list_.ptr;
    {
        const uint64_t ntiles = (uint64_t)div_up(V_, kThreads >> pull_gshift_) * (Sp_ == 1 ? 1 : (Sp_ / 4) >> pull_gshift_);
        auto gcd = [](uint64_t x, uint64_t y) { while (y) { const uint64_t t = x % y; x = y; y = t; } return x; };
        uint64_t k = std::max<uint64_t>(1, (uint64_t)(0.6180339887 * (double)ntiles)) | 1ull;
        while (gcd(k, ntiles) != 1) k += 2;
        a.pull_tile_mul = (uint32_t)(k % std::max<uint64_t>(ntiles, 1));
        if (a.pull_tile_mul == 0) a.pull_tile_mul = 1;
    }
    a.iterlog = iterlog_.ptr;
    a.iterlog_cap = iterlog_.ptr ? kIterLogCap : 0;
    a.ctalog = ctalog_.ptr;
    a.probe_iter = env_int("DPPR_PROBE_ITER", 10);
    if (mode_ == DPPR_ENGINE_STEPWISE) {
        launch_push_stepwise(a);
        return;
    }
    if (mode_ == DPPR_ENGINE_ASYNC) {
        launch_push_async(a);
        return;
    }
    void *params[] = {(void *)&a};
    void *kern = nullptr;
    switch (cfg_.variant) {
        case 0: kern = dense_ ? persistent_kernel<0, true>() : persistent_kernel<0>(); break;
        case 1: kern = persistent_kernel<1>(); break;
        case 2: kern = persistent_kernel<2>(); break;
        default: kern = persistent_kernel<3>(); break;
    }
    DPPR_CUDA(cudaLaunchCooperativeKernel(kern, dim3(coop_grid_[cfg_.variant]), dim3(kThreads), params, 0, st_));
    ++launch_counter();
}

void Engine::launch_push_async(PushArgs &b) {
    AsyncArgs a{};
    a.base = b;
    for (int i = 0; i < 2; ++i) {
        a.q[i].slots = q_[i].ptr;
        a.q[i].slot_ru = qr_[i].ptr;
        a.q[i].mask = ring_cap_ - 1;
        a.q[i].tail = async_ctr_.ptr + (3 * i + 0) * 16;
        a.q[i].head = async_ctr_.ptr + (3 * i + 1) * 16;
        a.q[i].done = async_ctr_.ptr + (3 * i + 2) * 16;
        a.q[i].fence = async_ctr_.ptr + (6 + 2 * i) * 16;
        a.q[i].theta0 = async_ctr_.ptr + (7 + 2 * i) * 16;
    }
    a.guard_slots = 0;
    a.dbg = iterlog_.ptr ? (unsigned long long *)ctalog_.ptr : nullptr;
    if (a.dbg) {
        DPPR_CUDA(cudaMemsetAsync(a.dbg, 0, 128, st_));
        DPPR_CUDA(cudaMemsetAsync(a.dbg + 10, 0xff, 8, st_));
        DPPR_CUDA(cudaMemsetAsync(a.dbg + 12, 0xff, 8, st_));
    }
    {
        const char *g = std::getenv("DPPR_CARRY_GAMMA"), *sc = std::getenv("DPPR_CARRY_SCALE");
        a.carry_gamma = g ? std::atof(g) : 0.7;
        a.carry_scale = sc ? std::atof(sc) : 0.01;
    }
    DPPR_CUDA(cudaMemsetAsync(async_ctr_.ptr, 0, async_ctr_.bytes(), st_));
    void *params[] = {(void *)&a};
    DPPR_CUDA(cudaLaunchCooperativeKernel((void *)push_async, dim3(async_grid_), dim3(kThreads), params, 0, st_));
    ++launch_counter();
}

// Debug / profiling mode with the reference's structure: one launch per sub-pass and a blocking
// read of the frontier counters per iteration (gpu/PPRRevPushGPU.cuh:106-108).
void Engine::launch_push_stepwise(PushArgs &a) {
    const int grid = coop_grid_[cfg_.variant];
    const int var = cfg_.variant;
    uint32_t it = 0;
    PushCtrl h{};
    const int nphases = a.init_mode ? 1 : 2;
    for (int phase = 0; phase < nphases; ++phase) {
        if (phase > 0) {  // same slot hygiene as push_persistent at a phase change
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
                std::memcpy(&t0, &h.theta0[phase], sizeof(double));
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
    for (auto &m : meta_)
        for (auto &ev : m.ev)
            if (ev) cudaEventDestroy(ev);
    meta_.clear();
    meta_.emplace_back();
    record(0);
    state_init<<<grid_for(Vp_ * S_), kThreads, 0, st_>>>(p_.ptr, r_.ptr, status_.ptr, Vp_, S_, src_.ptr); ++launch_counter();
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
    DPPR_CUDA(cudaSetDevice(dev_));
    meta_.emplace_back();
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
    DPPR_CUDA(cudaSetDevice(dev_));
    meta_.emplace_back();
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
    meta_.emplace_back();
    record(0);
    record(1);
    apply_batch_common((const int2 *)dpairs, B);
}

void Engine::apply_batch_common(const int2 *arriving, int64_t B) {
    BatchMeta &m = cur();
    const int64_t nA = 2 * D_ * B;  // entries in a group (directed: 2B per group, undirected: 4B in the one group)
    m.edges = B;
    m.entries = nA;
    m.has_window = true;
    int *werr = (int *)(counters_.ptr + 3);
    DPPR_CUDA(cudaMemsetAsync(counters_.ptr, 0, sizeof(uint32_t) * 3, st_));
    WindowView wv{V_, vmeta_.ptr, pool_.ptr, outdeg_.ptr, pool_top_.ptr, pool_cap_, werr};
    WindowView wvo{V_, vmeta_out_.ptr, pool_.ptr, outdeg_.ptr, pool_top_.ptr, pool_cap_, werr};  // vmeta null = no out-lists
    if (nA <= kFusedMaxEntries && env_int("DPPR_FUSED_WINDOW", 1)) {
        // small batch: the whole update in one single-CTA launch (window_fused.cuh)
        FusedArgs f{};
        f.log = log_.ptr; f.W = W_; f.log_start = log_start_; f.arriving = arriving; f.B = B;
        f.directed = D_ == 1; f.key_bits = key_bits_;
        for (int i = 0; i < 2; ++i) { f.akey[i] = akey_[i].ptr; f.aval[i] = aval_[i].ptr; f.bkey[i] = bkey_[i].ptr; f.bval[i] = bval_[i].ptr; }
        f.segA = segA_; f.segB = segB_; f.w = wv;
        f.ins_pos = ins_pos_.ptr; f.jobs = jobs_.ptr; f.njobs = counters_.ptr + 2; f.seg_d0 = seg_d0_.ptr; f.perm = perm_.ptr;
        f.wo = wvo; f.ins_posB = ins_posB_.ptr; f.jobsB = jobsB_.ptr; f.njobsB = counters_.ptr + 5;
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
    if (nA <= kCoopMaxEntries && coop_win_grid_ > 0 && env_int("DPPR_COOP_WINDOW", 1)) {
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
                                                        akey_[0].ptr, aval_[0].ptr, bkey_[0].ptr, bval_[0].ptr, werr, perm_.ptr); ++launch_counter();
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
    dim3 g((unsigned)std::min(grid_for(n), 8 * sm_count_), (unsigned)S_);
    repair_accumulate<<<g, kThreads, 0, st_>>>(sb_val_, n, segB_.segof, p_.ptr, Vp_, delta_.ptr, Nb_); ++launch_counter();
    repair_finalize<<<grid_for(n * S_), kThreads, 0, st_>>>(segB_, seg_d0_.ptr, src_.ptr, S_, p_.ptr, r_.ptr, Vp_,
                                                          delta_.ptr, Nb_, cfg_.alpha); ++launch_counter();
    DPPR_CUDA(cudaGetLastError());
    record(3);
    if (!repair_only) launch_push(false);
    else DPPR_CUDA(cudaMemsetAsync(ctrl_.ptr, 0, kCtrlZeroBytes, st_));
    record(4);
    finish_record();
    batch_pending_ = false;
}

void Engine::sync() {
    DPPR_CUDA(cudaSetDevice(dev_));
    DPPR_CUDA(cudaStreamSynchronize(st_));
}

void Engine::get_stats(int64_t batch_index, dppr_batch_stats *out) {
    if (!out) throw InvalidArgument("null stats pointer");
    if (meta_.empty()) throw StateError("no batch has been processed yet");
    if (batch_index < 0) batch_index = (int64_t)meta_.size() - 1;
    if (batch_index >= (int64_t)meta_.size()) throw InvalidArgument("batch index out of range");
    if (batch_index == (int64_t)meta_.size() - 1 && batch_pending_) throw StateError("batch applied but not refreshed yet");
    sync();
    const BatchMeta &m = meta_[(size_t)batch_index];
    const BatchRecord *rec = record_slot((size_t)batch_index);
    std::memset(out, 0, sizeof(*out));
    out->batch_index = batch_index;
    out->edges = m.edges;
    out->batch_entries = m.entries;  // N_b = 2*D*B
    out->touched_vertices = (D_ == 1) ? rec->nseg_out : rec->nseg_in;
    out->iterations = (int64_t)rec->ctrl.iters;
    out->frontier_pops = (int64_t)rec->ctrl.pops - (int64_t)rec->ctrl.carried;  // carried items are not pushed
    out->traversed_edges = (int64_t)rec->ctrl.edges;
    out->hub_pops = (int64_t)rec->ctrl.hubs;
    out->relocations = rec->njobs;
    out->pool_used = (int64_t)rec->pool_top;
    out->error_flags = rec->ctrl.errflags;
    out->dense_sweeps = (int32_t)rec->ctrl.sweeps;
    auto ms = [&](int a, int b) -> float {
        if (!cfg_.record_timing || !m.ev[a] || !m.ev[b]) return 0.f;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, m.ev[a], m.ev[b]) != cudaSuccess) return 0.f;
        return t;
    };
    out->ms_upload = m.has_upload ? ms(0, 1) : 0.f;
    out->ms_window = m.has_window ? ms(1, 2) : 0.f;
    out->ms_repair = m.has_window ? ms(2, 3) : 0.f;
    out->ms_push = ms(3, 4);
}

void Engine::get_vector(int which, int32_t s, double *out) {
    if (!out) throw InvalidArgument("null output pointer");
    if (s < 0 || s >= S_) throw InvalidArgument("source index out of range");
    if (!solved_) throw StateError("no estimates before dppr_solve_initial");
    sync();
    const double *src = (which == 0 ? p_.ptr : r_.ptr) + (size_t)s * Vp_;
    if (!perm_.ptr) {
        DPPR_CUDA(cudaMemcpy(out, src, sizeof(double) * (size_t)V_, cudaMemcpyDeviceToHost));
        return;
    }
    DevBuf<double> tmp;
    tmp.alloc((size_t)V_);
    gather_by_perm<double><<<grid_for(V_), kThreads, 0, st_>>>(src, 1, perm_.ptr, tmp.ptr, V_); ++launch_counter();
    DPPR_CUDA(cudaStreamSynchronize(st_));
    DPPR_CUDA(cudaMemcpy(out, tmp.ptr, sizeof(double) * (size_t)V_, cudaMemcpyDeviceToHost));
}

void Engine::copy_estimates_device(int32_t s, void *dptr) {
    if (!dptr) throw InvalidArgument("null device pointer");
    if (s < 0 || s >= S_) throw InvalidArgument("source index out of range");
    DPPR_CUDA(cudaSetDevice(dev_));
    gather_by_perm<double><<<grid_for(V_), kThreads, 0, st_>>>(p_.ptr + (size_t)s * Vp_, 1, perm_.ptr, (double *)dptr, V_); ++launch_counter();
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
        double *dst = (which == 0 ? p_.ptr : r_.ptr) + (size_t)s * Vp_;
        DPPR_CUDA(cudaMemcpy(tmp.ptr, h, sizeof(double) * (size_t)V_, cudaMemcpyHostToDevice));
        scatter_by_perm<double><<<grid_for(V_), kThreads, 0, st_>>>(tmp.ptr, perm_.ptr, dst, V_); ++launch_counter();
        DPPR_CUDA(cudaStreamSynchronize(st_));
    }
    solved_ = true;
    if (meta_.empty()) meta_.emplace_back();
}

int Engine::get_iterlog(uint32_t *out, int cap) {
    if (!iterlog_.ptr || meta_.empty()) return 0;
    sync();
    const BatchRecord *rec = record_slot(meta_.size() - 1);
    int n = (int)std::min<unsigned long long>(rec->ctrl.iters, (unsigned long long)std::min(cap, kIterLogCap));
    if (n > 0) DPPR_CUDA(cudaMemcpy(out, iterlog_.ptr, sizeof(uint4) * (size_t)n, cudaMemcpyDeviceToHost));
    return n;
}

int Engine::get_ctalog(unsigned long long *out, int cap_rows) {
    if (!ctalog_.ptr) return 0;
    sync();
    int rows = std::min(cap_rows, coop_grid_[cfg_.variant]);
    if (use_async()) rows = 2;
    DPPR_CUDA(cudaMemcpy(out, ctalog_.ptr, sizeof(unsigned long long) * 8 * (size_t)rows, cudaMemcpyDeviceToHost));
    return rows;
}

// canonical CSR: rows ascending, duplicates kept.  Device sort, test/validation path.
void Engine::export_csr(int32_t *in_row_ptr, int32_t *in_col_ind, int32_t *out_deg, bool out_lists) {
    if (!window_ready_) throw StateError("dppr_export_window_csr before dppr_init_window");
    if (out_lists && !outlists_ && D_ == 1) throw StateError("this engine does not maintain out-lists (dense iterations are off)");
    const uint4 *vm = (out_lists && outlists_) ? vmeta_out_.ptr : vmeta_.ptr;
    sync();
    DevBuf<uint32_t> len, rowptr, key[2], val[2], scratch, total;
    len.alloc((size_t)V_); rowptr.alloc((size_t)V_ + 1);
    for (int i = 0; i < 2; ++i) { key[i].alloc((size_t)Ew_); val[i].alloc((size_t)Ew_); }
    scratch.alloc(std::max(sort_scratch_elems(Ew_), scan_scratch_elems(V_)));
    total.alloc(1);
    win_export_len<<<grid_for(V_), kThreads, 0, st_>>>(vm, len.ptr, V_, perm_.ptr); ++launch_counter();
    exclusive_scan<uint32_t>(len.ptr, rowptr.ptr, V_, scratch.ptr, total.ptr, st_);
    DPPR_CUDA(cudaMemcpyAsync(rowptr.ptr + V_, total.ptr, sizeof(uint32_t), cudaMemcpyDeviceToDevice, st_));
    uint32_t htotal = 0;
    DPPR_CUDA(cudaMemcpyAsync(&htotal, total.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st_));
    DPPR_CUDA(cudaStreamSynchronize(st_));
    if ((int64_t)htotal != Ew_)
        throw StateError("window graph holds " + std::to_string(htotal) + " entries, expected " + std::to_string(Ew_));
    win_export_entries<<<grid_for((int64_t)V_ * 32), kThreads, 0, st_>>>(vm, pool_.ptr, rowptr.ptr, key[0].ptr,
                                                                      val[0].ptr, V_, perm_.ptr, inv_.ptr); ++launch_counter();
    // sort by (dst, src): LSD over the pair = stable sort by src, then stable sort by dst
    int res = sort_pairs(val[0].ptr, key[0].ptr, val[1].ptr, key[1].ptr, Ew_, key_bits_, scratch.ptr, st_);
    uint32_t *k0 = key[res].ptr, *v0 = val[res].ptr, *k1 = key[1 - res].ptr, *v1 = val[1 - res].ptr;
    res = sort_pairs(k0, v0, k1, v1, Ew_, key_bits_, scratch.ptr, st_);
    const uint32_t *cols = res ? v1 : v0;
    DPPR_CUDA(cudaGetLastError());
    DPPR_CUDA(cudaStreamSynchronize(st_));
    if (in_row_ptr) DPPR_CUDA(cudaMemcpy(in_row_ptr, rowptr.ptr, sizeof(int32_t) * ((size_t)V_ + 1), cudaMemcpyDeviceToHost));
    if (in_col_ind && Ew_ > 0) DPPR_CUDA(cudaMemcpy(in_col_ind, cols, sizeof(int32_t) * (size_t)Ew_, cudaMemcpyDeviceToHost));
    if (out_deg) {
        DevBuf<int32_t> od;
        od.alloc((size_t)V_);
        gather_by_perm<int32_t><<<grid_for(V_), kThreads, 0, st_>>>(outdeg_.ptr, 1, perm_.ptr, od.ptr, V_); ++launch_counter();
        DPPR_CUDA(cudaStreamSynchronize(st_));
        DPPR_CUDA(cudaMemcpy(out_deg, od.ptr, sizeof(int32_t) * (size_t)V_, cudaMemcpyDeviceToHost));
    }
}

}  // namespace dppr
