// engine.cuh -- host-side owner of the device state and per-batch orchestration.
//
// Plays the role of DeviceMemory (gpu/DeviceMemory.cuh), SlidingGraphBuilder
// (gpu/SlidingGraphBuilder.cuh) and the PPRGPU / PPRRevPushGPU* drivers (gpu/PPRGPU.cuh,
// gpu/PPRRevPushGPU.cuh, gpu/PPRRevPushGPUVariants.cuh) of the reference, behind include/dppr.h.
// One engine = one GPU = one stream; nothing here blocks the host except the explicit sync /
// read-back entry points (and the stepwise debug mode).
#pragma once
#include <vector>
#include <string>
#include <memory>
#include "../../include/dppr.h"
#include "common.cuh"
#include "primitives.cuh"
#include "window.cuh"
#include "repair.cuh"
#include "window_fused.cuh"
#include "window_coop.cuh"
#include "push.cuh"
#include "topk.cuh"
#include "validate.cuh"

namespace dppr {

// what one refresh leaves behind for the host (copied D2H asynchronously into pinned memory)
struct BatchRecord {
    unsigned long long iters, pops, edges, gath, hubs, carried, dpops;
    unsigned long long walk_slots, walk_pairs, units;  // dense sweeps: out-list entries walked, (entry, source) gathers, (vertex, source) units
    unsigned long long pool_top, pool_leaked;
    uint32_t sweeps, nseg_in, nseg_out, njobs;
    int errflags;
    int arrived;  // 1 once the copy has landed (the host clears the slot before the batch is enqueued)
};

// every knob, resolved ONCE in the constructor: dppr_config::tuning, else the DPPR_* environment (development aid), else default
struct Tuning {
    bool relabel = true, relabel_both = false;
    int relabel_blocks = 1024;
    int ctas_per_sm = 4, tile_cap = 128, max_iters = 400000;
    int dense = 0;  // 0 auto, 1 always the switching kernel, -1 scatter only
    int dense_accel = 0;      // 0 / 1: Chebyshev-accelerated sweeps on undirected windows, -1: off
    int signed_push = 0;      // 0 / 1: variant 0 pushes both signs in one pass, -1: the reference's two passes
    double accel_frac = 0.9;  // ... while the frontier holds at least this fraction of all (vertex, source) pairs
    double dense_div = 4.0, dense_min_edges = -1.0;  // (< 0: 2e7 for variant 0, 0 for variants 1-3)
    int pull_group = 16, pull_warp_min = 32, pull_big_min = 0, pull_big_chunk = 0;  // (0: by the number of sources)
    double carry_gamma = 1.0, carry_scale = 0.01;
    int window_path = 0;  // 0 auto, 1 multi-kernel only, 2 cooperative or multi-kernel (no single-CTA kernel)
    bool iterlog = false;
    int probe_iter = 10;
    int panel_sources = 128;  // most sources per launch (see Engine::Pw_)
    int pull_warp_units = 8;  // multi-source sweeps: warp-slices per work item handed to a warp; 0: items go to CTAs
};

class Engine {
public:
    explicit Engine(const dppr_config &cfg);
    ~Engine();

    void init_window_pairs(const int32_t *pairs, int64_t n);
    void init_window_soa(const int32_t *e1, const int32_t *e2, int64_t n);
    void init_window_device_pairs(const int32_t *dpairs, int64_t n);
    void solve_initial();
    void apply_batch_host_pairs(const int32_t *pairs, int64_t B);
    void apply_batch_host_soa(const int32_t *e1, const int32_t *e2, int64_t B);
    void apply_batch_device_pairs(const int32_t *dpairs, int64_t B);
    void refresh(bool repair_only = false);
    void sync();
    void wait_event(void *cuda_event);
    void get_stats(int64_t batch_index, dppr_batch_stats *out);
    int64_t batches_done() const { return (int64_t)meta_.size() - 1; }
    void get_vector(int which, int32_t source_index, double *out);  // 0 = p, 1 = r
    void copy_estimates_device(int32_t source_index, void *dptr);
    void set_state(int32_t source_index, const double *p, const double *r);
    void export_csr(int32_t *in_row_ptr, int32_t *in_col_ind, int32_t *out_deg, bool out_lists = false);
    bool has_out_lists() const { return outlists_; }
    int64_t csr_entries() const { return Ew_; }
    void topk(int32_t first_source, int32_t n_sources, int32_t k, int32_t *ids, double *values);
    void validate(int32_t source_index, double *max_abs_residual, double *max_invariant_defect);
    void check_window_device(const int32_t *dpairs, int64_t n, int64_t *mismatches);
    int get_iterlog(uint32_t *out, int cap);
    int get_ctalog(unsigned long long *out, int cap_rows);  // debug: 8 stamps per CTA of the probed iteration

    std::string last_error;

private:
    static constexpr int kEventSlots = 64;
    struct BatchMeta {
        int64_t edges = 0, entries = 0;
        bool has_upload = false, has_window = false, resolved = false;
        float ms[4] = {0.f, 0.f, 0.f, 0.f};  // upload, window, repair, push
    };
    struct EventSet {
        cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // start, uploaded, window, repair, push
        bool used[5] = {false, false, false, false, false};
        int64_t batch = -1;
    };
    void begin_batch(int64_t edges, int64_t entries);
    void check_health();          // throws if a finished batch left device error flags (sticky)
    void resolve_events(EventSet &es);
    void apply_batch_common(const int2 *arriving, int64_t B);
    void build_initial_window();
    void launch_push(bool init_mode);
    void launch_push_panel(bool init_mode, int panel);
    void launch_push_stepwise(PushArgs &a);
    void record(int which);
    void finish_record();
    BatchMeta &cur() { return meta_.back(); }
    BatchRecord *record_slot(size_t k);
    int2 *stage_pairs(const int32_t *pairs, const int32_t *e1, const int32_t *e2, int64_t n);
    void validate_host_ids(const int32_t *pairs, const int32_t *e1, const int32_t *e2, int64_t n) const;
    int grid_for(int64_t n) const;
    PoolFree pool_free_view(int parity);
    // canonical (dst, src)-sorted entries of the window graph in device memory; returns pointers into the buffers
    struct SortedCsr {
        DevBuf<uint32_t> len, rowptr, key[2], val[2], scratch, total;
        const uint32_t *rows = nullptr, *cols = nullptr;
    };
    void export_sorted_device(SortedCsr &out, bool out_lists);

    dppr_config cfg_;
    Tuning tn_;
    std::vector<int32_t> sources_;
    int dev_ = 0, sm_count_ = 0, coop_grid_[4] = {0, 0, 0, 0}, mode_ = 0;
    cudaStream_t st_ = nullptr;
    int32_t V_ = 0;
    int64_t Sr_ = 1, W_ = 0, Ew_ = 0, Bmax_ = 0, Nb_ = 0;  // Sr_: row stride of the vertex-major state (1, or Pw_ rounded up to 8)
    int D_ = 1, S_ = 1, key_bits_ = 1;
    // Source panels: the S_ sources are split into n_panels_ equal groups of at most Pw_; every per-(vertex, source) array
    // holds one contiguous [V][Sr_] block per panel, and a refresh runs the panels one after the other (repair and push
    // launches take a panel's block as if it were the whole state).  Measured on BASELINE configs[3], 1000 sources on one
    // GPU: in ONE launch the gathered x rows are 2 KB apart pieces of a 6 GB array, the episode lasts until the slowest
    // source has thinned out (22 sweeps instead of 19) and DRAM traffic per source is 1.4x that of a 125-source launch.
    int Pw_ = 1, n_panels_ = 1;
    int64_t panel_stride_ = 0;                                       // elements between panels = V_ * Sr_
    int panel_sources(int k) const { return std::min(Pw_, S_ - k * Pw_); }
    size_t elem_base(int s) const { return (size_t)(s / Pw_) * (size_t)panel_stride_ + (size_t)(s % Pw_); }  // element (vertex 0, source s)
    bool window_ready_ = false, solved_ = false, batch_pending_ = false, debug_ = false;
    int64_t log_start_ = 0;
    int step_level_ = 0;  // stepwise mode keeps the status level on the host
    int failed_code_ = 0;         // sticky: DPPR_E_* once a batch left device error flags
    std::string failed_msg_;
    size_t health_checked_ = 0;   // records [0, health_checked_) have been inspected

    // window
    DevBuf<int2> log_;
    DevBuf<uint4> vmeta_;
    DevBuf<int32_t> pool_, outdeg_;
    DevBuf<unsigned long long> pool_top_;   // [0] bump pointer, [1] leaked slots
    DevBuf<uint32_t> perm_, inv_;   // internal vertex order (empty = identity)
    // free ranges of the pool (window.cuh, PoolFree)
    DevBuf<int> fcount_;
    DevBuf<uint32_t> foff_, fstack_, npend_;
    DevBuf<uint2> pend_[2];
    int64_t win_batches_ = 0;       // window updates so far (parity of the pending-release lists)
    // dense iterations in gather form (pull.cuh)
    DevBuf<uint4> vmeta_out_;       // out-lists of a directed graph (undirected: the in-lists serve)
    DevBuf<uint32_t> ins_posB_;
    DevBuf<RelocJob> jobsB_;
    DevBuf<uint16_t> x_[2];
    DevBuf<double> bigacc_;
    DevBuf<HubItem> big_;
    DevBuf<uint32_t> tile_list_;
    uint32_t bigcap_ = 0, tile_cap_ = 0;
    size_t dyn_smem_ = 0;
    int pull_gshift_ = 0, pull_big_min_ = 0, pull_big_chunk_ = 0;
    bool dense_ = false, outlists_ = false;
    unsigned long long pool_cap_ = 0;
    // batch scratch
    DevBuf<int2> arriving_;
    DevBuf<uint32_t> akey_[2], aval_[2], bkey_[2], bval_[2];
    DevBuf<uint32_t> sort_scratch_, flags_;
    DevBuf<uint32_t> segA_vertex_, segA_start_, segA_first_, segA_of_, segB_vertex_, segB_start_, segB_first_, segB_of_;
    DevBuf<uint32_t> counters_;  // [0]=nsegA [1]=nsegB [2]=njobs [3]=window error flags [4]=coop barrier [5]=njobsB
    DevBuf<uint32_t> tile_heads_;
    int coop_win_grid_ = 0;
    DevBuf<uint32_t> ins_pos_;
    DevBuf<RelocJob> jobs_;
    DevBuf<int32_t> seg_d0_;
    DevBuf<double> delta_;
    // sorted views of the current batch
    uint32_t *sa_key_ = nullptr, *sa_val_ = nullptr, *sb_key_ = nullptr, *sb_val_ = nullptr;
    Segments segA_{}, segB_{};
    // state
    DevBuf<double> p_, r_;
    DevBuf<int32_t> status_, src_;
    // push
    DevBuf<unsigned long long> q_[2], qalt_;
    DevBuf<double> qr_[2];
    DevBuf<HubItem> hub_[2];
    uint32_t qcap_ = 0, hcap_ = 0;
    DevBuf<float> pacc_;                 // per (vertex, source): amounts popped by the running sweep episode (several sources)
    DevBuf<PushCtrl> ctrl_, ctrl_acc_;   // ctrl_acc_: counters of a refresh summed over its panels
    DevBuf<BatchRecord> dev_record_;
    DevBuf<uint4> iterlog_;
    DevBuf<unsigned long long> ctalog_;
    static constexpr int kIterLogCap = 4096;
    // top-k read-back
    DevBuf<unsigned long long> topk_key_;
    DevBuf<uint32_t> topk_id_;
    DevBuf<int32_t> topk_out_ids_;
    DevBuf<double> topk_out_vals_;
    PinnedBuf<int32_t> topk_host_ids_;
    PinnedBuf<double> topk_host_vals_;
    DevBuf<uint32_t> topk_prev_, topk_cid_;       // fast path: previous winners (internal ids), candidate ids
    DevBuf<int> topk_prevk_, topk_over_;
    DevBuf<unsigned long long> topk_bound_, topk_ckey_;
    DevBuf<unsigned int> topk_count_;
    PinnedBuf<int> topk_host_over_;
    // host staging
    static constexpr int kStageSlots = 4;
    PinnedBuf<int2> hstage_[kStageSlots];
    cudaEvent_t hstage_free_[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    int hstage_next_ = 0;
    std::vector<std::unique_ptr<PinnedBuf<BatchRecord>>> records_;
    static constexpr size_t kRecordsPerChunk = 1024;
    std::vector<BatchMeta> meta_;
    EventSet events_[kEventSlots];
};

}  // namespace dppr
