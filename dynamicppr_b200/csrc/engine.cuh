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
#include "push_async.cuh"

namespace dppr {

// what one refresh leaves behind for the host (copied D2H asynchronously into pinned memory)
struct BatchRecord {
    PushCtrl ctrl;
    uint32_t nseg_in, nseg_out, njobs, pad;
    unsigned long long pool_top;
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
    void get_stats(int64_t batch_index, dppr_batch_stats *out);
    int64_t batches_done() const { return (int64_t)meta_.size() - 1; }
    void get_vector(int which, int32_t source_index, double *out);  // 0 = p, 1 = r
    void copy_estimates_device(int32_t source_index, void *dptr);
    void set_state(int32_t source_index, const double *p, const double *r);
    void export_csr(int32_t *in_row_ptr, int32_t *in_col_ind, int32_t *out_deg, bool out_lists = false);
    bool has_out_lists() const { return outlists_; }
    int64_t csr_entries() const { return Ew_; }
    int get_iterlog(uint32_t *out, int cap);
    int get_ctalog(unsigned long long *out, int cap_rows);  // debug: 8 stamps per CTA of the probed iteration  // debug: 4 uint32 per iteration of the last refresh

    std::string last_error;

private:
    struct BatchMeta {
        int64_t edges = 0, entries = 0;
        cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // start, uploaded, window, repair, push
        bool has_upload = false, has_window = false;
    };
    void apply_batch_common(const int2 *arriving, int64_t B);
    void build_initial_window();
    void launch_push(bool init_mode);
    void launch_push_stepwise(PushArgs &a);
    void launch_push_async(PushArgs &a);
    bool use_async() const { return mode_ == DPPR_ENGINE_ASYNC; }
    void record(int which);
    void finish_record();
    BatchMeta &cur() { return meta_.back(); }
    BatchRecord *record_slot(size_t k);
    int2 *stage_pairs(const int32_t *pairs, const int32_t *e1, const int32_t *e2, int64_t n);
    int grid_for(int64_t n) const;

    dppr_config cfg_;
    std::vector<int32_t> sources_;
    int dev_ = 0, sm_count_ = 0, coop_grid_[4] = {0, 0, 0, 0}, async_grid_ = 0, mode_ = 0;
    cudaStream_t st_ = nullptr;
    int32_t V_ = 0;
    int64_t Vp_ = 0, W_ = 0, Ew_ = 0, Bmax_ = 0, Nb_ = 0;
    int D_ = 1, S_ = 1, key_bits_ = 1;
    bool window_ready_ = false, solved_ = false, batch_pending_ = false;
    int64_t log_start_ = 0;
    int step_level_ = 0;  // stepwise mode keeps the status level on the host

    // window
    DevBuf<int2> log_;
    DevBuf<uint4> vmeta_;
    DevBuf<int32_t> pool_, outdeg_;
    DevBuf<unsigned long long> pool_top_;
    DevBuf<uint32_t> perm_, inv_;   // internal vertex order (empty = identity)
    bool relabel_ = true;
    // dense iterations in gather form (pull.cuh)
    DevBuf<uint4> vmeta_out_;       // out-lists of a directed graph (undirected: the in-lists serve)
    DevBuf<uint32_t> ins_posB_;
    DevBuf<RelocJob> jobsB_;
    DevBuf<double> x_[2], bigacc_;
    DevBuf<HubItem> big_;
    DevBuf<uint32_t> tile_list_;
    uint32_t bigcap_ = 0;
    int Sp_ = 1, pull_gshift_ = 0;
    bool dense_ = false, outlists_ = false;
    double dense_div_ = 0.0;
    int pull_warp_min_ = 0, pull_cta_min_ = 0, pull_big_min_ = 0;
    unsigned long long pool_cap_ = 0;
    // batch scratch
    DevBuf<int2> arriving_;
    DevBuf<uint32_t> akey_[2], aval_[2], bkey_[2], bval_[2];
    DevBuf<uint32_t> sort_scratch_, flags_;
    DevBuf<uint32_t> segA_vertex_, segA_start_, segA_first_, segA_of_, segB_vertex_, segB_start_, segB_first_, segB_of_;
    DevBuf<uint32_t> counters_;  // [0]=nsegA [1]=nsegB [2]=njobs [3]=window error flags [4]=coop barrier
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
    DevBuf<unsigned long long> q_[2];
    DevBuf<double> qr_[2];
    DevBuf<HubItem> hub_[2];
    uint32_t qcap_ = 0, hcap_ = 0;
    unsigned long long ring_cap_ = 0;           // async: slots per ring (power of two)
    DevBuf<unsigned long long> async_ctr_;      // async: tail/head/done of both rings, one 128-byte line each
    int guard_slots_ = 0;
    DevBuf<PushCtrl> ctrl_;
    DevBuf<BatchRecord> dev_record_;
    DevBuf<uint4> iterlog_;
    DevBuf<unsigned long long> ctalog_;
    static constexpr int kIterLogCap = 4096;
    // host staging
    static constexpr int kStageSlots = 4;
    PinnedBuf<int2> hstage_[kStageSlots];
    cudaEvent_t hstage_free_[kStageSlots] = {nullptr, nullptr, nullptr, nullptr};
    int hstage_next_ = 0;
    std::vector<std::unique_ptr<PinnedBuf<BatchRecord>>> records_;
    static constexpr size_t kRecordsPerChunk = 1024;
    std::vector<BatchMeta> meta_;
};

}  // namespace dppr
