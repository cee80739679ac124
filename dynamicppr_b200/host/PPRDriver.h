// PPRDriver.h -- host orchestration of one run, the role of PPRGPU::DynamicExecute /
// SlidingWindowExecuteMainLoop (gpu/PPRGPU.cuh:64-177) and of the variant classes chosen in
// gpu/PPRGPUMain.cu:24-27.  Everything below the C ABI (include/dppr.h) is CUDA; this file only
// feeds batches and prints the keys the reference's scripts scrape (scripts/extract_gpu.py:18-32):
// the LAST line containing `ppr_latency` / `ppr_throughput`, value = token after the first space.
// Extra keys therefore never contain those two substrings.
#pragma once
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "dppr.h"
#include "Arguments.h"
#include "EdgeStream.h"
#include "SourcePicker.h"

namespace dppr_host {

inline void die_on(int rc, dppr_engine *e, const char *what) {
    if (rc == DPPR_OK) return;
    // the reference's convention for device failures: print and exit(-1) (gpu/GPUUtil.cuh:7-19)
    std::cout << what << " failed: " << dppr_last_error(e) << " (code " << rc << ")" << std::endl;
    std::exit(-1);
}

class PPRDriver {
public:
    PPRDriver(const Settings &s, EdgeStream &stream) : s_(s), stream_(stream) {
        sources_.push_back(s.source);
        if (!s.pick.empty()) {
            // the reference's protocol: `workload` writes ten ids per bucket, scripts/gpu.sh:17 runs the 4th (SOURCES_START=3)
            int64_t st, ed;
            bucket_ranks(s.pick, st, ed);
            const DegreeRanking r = rank_by_degree(s.device, stream.vertex_count, stream.directed, true, stream.all_pairs(),
                                                   (int64_t)stream.stream_length);
            const std::vector<int32_t> ids = choose_degree_range(r, 10, st, ed);
            sources_[0] = ids[3];
        }
        if (!s.sources_file.empty()) {
            sources_.clear();
            std::ifstream in(s.sources_file.c_str());
            int v;
            while (in >> v) sources_.push_back(v);
            if (sources_.empty()) {
                std::cout << "no source vertex in " << s.sources_file << std::endl;
                std::exit(-1);
            }
        }
        std::cout << "choose " << sources_[0] << " as source vertex id";
        if (sources_.size() > 1) std::cout << " (+" << sources_.size() - 1 << " more)";
        std::cout << std::endl;
        dppr_config c{};
        c.vertex_count = stream.vertex_count;
        c.directed = stream.directed ? 1 : 0;
        c.window_edges = stream.window;
        c.max_batch_edges = (int64_t)stream.per_batch;
        c.alpha = kAlpha;
        c.epsilon = s.tolerance;
        c.variant = s.variant;
        c.device = s.device;
        c.n_sources = (int32_t)sources_.size();
        c.sources = sources_.data();
        c.engine_mode = s.stepwise ? DPPR_ENGINE_STEPWISE : DPPR_ENGINE_AUTO;
        c.record_timing = 1;
        c.pool_factor = s.pool_factor;
        std::cout << "init sliding graph.." << std::endl;
        int rc = dppr_create(&c, &eng_);
        if (rc != DPPR_OK) {
            std::cout << "dppr_create failed: " << dppr_last_error(nullptr) << " (code " << rc << ")" << std::endl;
            std::exit(-1);
        }
        std::cout << "initialize gpu graph..." << std::endl;
        die_on(dppr_init_window_pairs(eng_, stream.initial_window(), stream.window), eng_, "dppr_init_window_pairs");
    }
    ~PPRDriver() { dppr_destroy(eng_); }

    // static graph: Init + push phase 0 (PPRGPU::Execute, gpu/PPRGPU.cuh:43-61)
    void Execute() {
        std::cout << "start Execute..." << std::endl;
        die_on(dppr_solve_initial(eng_), eng_, "dppr_solve_initial");
        dppr_batch_stats st;
        die_on(dppr_get_batch_stats(eng_, 0, &st), eng_, "dppr_get_batch_stats");
        check_health(st);
        std::cout << "elapsed time=" << st.ms_push << "ms" << std::endl;
        print_counters(st);
        if (s_.validate) validate(0, stream_.initial_window());
        dump();
    }

    // sliding window: initial solve, then one slide per batch (PPRGPU::DynamicExecute)
    void DynamicExecute() {
        std::cout << "start..." << std::endl;
        die_on(dppr_solve_initial(eng_), eng_, "dppr_solve_initial");
        if (s_.validate) validate(0, stream_.initial_window());
        const size_t B = stream_.per_batch;
        size_t k = 0;            // batches enqueued
        size_t reported = 0;     // batches folded into ppr_time
        while (k < stream_.batch_count) {
            // progress at the reference's cadence (gpu/PPRGPU.cuh:116): before EVERY batch when a batch holds more than 100
            // edges, else before every 100th (--progress 0: at the end only; the per-batch form waits for the device each time)
            if (s_.progress && (B > 100 || (k + 1) % 100 == 0)) {
                fold_stats(reported, k);
                print_keys(k + 1, k);
            }
            const int32_t *batch = stream_.next_batch();
            if (!batch) break;  // fewer than B edges remain: stop, as the reference does
            die_on(dppr_slide_pairs(eng_, batch, (int64_t)B), eng_, "dppr_slide_pairs");
            ++k;
            if (s_.validate) validate(k, stream_.window_after(k));
        }
        fold_stats(reported, k);
        std::cout << "finish!" << std::endl;
        print_extras(k);
        print_keys(k + 1, k);
        dump();
    }

private:
    void check_health(const dppr_batch_stats &st) {
        if (st.error_flags) {
            std::cout << "device error flags " << st.error_flags << " in batch " << st.batch_index
                      << " (1 pool, 2 frontier queue, 4 hub list, 8 watchdog, 16 underflow, 32 bad id)" << std::endl;
            std::exit(-1);
        }
    }
    void fold_stats(size_t &reported, size_t upto) {
        for (size_t b = reported + 1; b <= upto; ++b) {
            dppr_batch_stats st;
            die_on(dppr_get_batch_stats(eng_, (int64_t)b, &st), eng_, "dppr_get_batch_stats");
            check_health(st);
            const double ppr = (double)st.ms_repair + (double)st.ms_push;
            ppr_time_ += ppr;
            e2e_time_ += ppr + st.ms_upload + st.ms_window;
            ppr_samples_.push_back(ppr);
            e2e_samples_.push_back(ppr + st.ms_upload + st.ms_window);
            iters_ += st.iterations; pops_ += st.frontier_pops; edges_ += st.traversed_edges;
            relocs_ += st.relocations; pool_used_ = st.pool_used;
            if (s_.stats)
                std::cout << "batch " << b << " iterations=" << st.iterations << " pops=" << st.frontier_pops
                          << " traversed=" << st.traversed_edges << " window_ms=" << st.ms_window
                          << " repair_ms=" << st.ms_repair << " push_ms=" << st.ms_push << std::endl;
        }
        reported = upto;
    }
    static double pct(std::vector<double> v, double q) {
        if (v.empty()) return 0;
        std::sort(v.begin(), v.end());
        size_t i = (size_t)(q * (double)(v.size() - 1) + 0.5);
        return v[i];
    }
    // the four keys of gpu/PPRGPU.cuh:170-176, same order and formulas
    void print_keys(size_t stream_batch_count, size_t done) {
        std::cout << "coming stream_batch_count=" << stream_batch_count << std::endl;
        long long cur_edge_num = (long long)stream_.per_batch * (long long)done;
        std::cout << "ppr_time " << ppr_time_ << std::endl;
        std::cout << "edge_num " << cur_edge_num << std::endl;
        std::cout << "ppr_latency " << (done > 0 ? ppr_time_ / (double)done : 0) << std::endl;
        std::cout << "ppr_throughput " << (double)cur_edge_num / ppr_time_ * 1000.0 << std::endl;
    }
    void print_extras(size_t done) {
        std::cout << "batches_done " << done << std::endl;
        std::cout << "sources " << sources_.size() << std::endl;
        std::cout << "source_batches_per_s " << (e2e_time_ > 0 ? (double)sources_.size() * (double)done / e2e_time_ * 1000.0 : 0) << std::endl;
        std::cout << "p50_batch_ms " << pct(ppr_samples_, 0.5) << std::endl;
        std::cout << "p95_batch_ms " << pct(ppr_samples_, 0.95) << std::endl;
        std::cout << "e2e_time_ms " << e2e_time_ << std::endl;
        std::cout << "e2e_p50_batch_ms " << pct(e2e_samples_, 0.5) << std::endl;
        std::cout << "e2e_edges_per_s " << (e2e_time_ > 0 ? (double)stream_.per_batch * (double)done / e2e_time_ * 1000.0 : 0) << std::endl;
        std::cout << "push_iterations " << iters_ << std::endl;
        std::cout << "frontier_pops " << pops_ << std::endl;
        std::cout << "traversed_edges " << edges_ << std::endl;
        std::cout << "ring_relocations " << relocs_ << std::endl;
        std::cout << "pool_slots_used " << pool_used_ << std::endl;
    }
    // the reference's -DVALIDATE (gpu/PPRGPU.cuh:58-60,90-92,165-167), on the device: window graph bit-exact against the
    // file's window, |r| within eps, push invariant (implies |p - pi| <= eps) -- exit(-1) on a violation, like its asserts
    void validate(size_t batch, const int32_t *window_pairs) {
        int64_t bad = -1;
        die_on(dppr_check_window(eng_, window_pairs, stream_.window, &bad), eng_, "dppr_check_window");
        double worst_r = 0, worst_inv = 0;
        for (size_t i = 0; i < sources_.size(); ++i) {
            double mr = 0, inv = 0;
            die_on(dppr_validate(eng_, (int32_t)i, &mr, &inv), eng_, "dppr_validate");
            worst_r = std::max(worst_r, mr);
            worst_inv = std::max(worst_inv, inv);
        }
        std::cout << "validate batch " << batch << ": window_mismatches=" << bad << " max_abs_residual=" << worst_r
                  << " invariant_defect=" << worst_inv << std::endl;
        if (bad != 0 || !(worst_r <= s_.tolerance) || !(worst_inv <= 1e-12)) {
            std::cout << "VALIDATION FAILED" << std::endl;
            std::exit(-1);
        }
    }
    void print_counters(const dppr_batch_stats &st) {
        std::cout << "push_iterations " << st.iterations << std::endl;
        std::cout << "frontier_pops " << st.frontier_pops << std::endl;
        std::cout << "traversed_edges " << st.traversed_edges << std::endl;
    }
    void dump() {
        if (s_.dump_file.empty()) return;
        FILE *f = std::fopen(s_.dump_file.c_str(), "wb");
        if (!f) {
            std::cout << "cannot write " << s_.dump_file << std::endl;
            std::exit(-1);
        }
        std::vector<double> p((size_t)stream_.vertex_count);
        for (size_t i = 0; i < sources_.size(); ++i) {
            die_on(dppr_get_estimates(eng_, (int32_t)i, p.data()), eng_, "dppr_get_estimates");
            std::fwrite(p.data(), sizeof(double), p.size(), f);
        }
        std::fclose(f);
    }

    Settings s_;
    EdgeStream &stream_;
    std::vector<int32_t> sources_;
    dppr_engine *eng_ = nullptr;
    double ppr_time_ = 0, e2e_time_ = 0;
    std::vector<double> ppr_samples_, e2e_samples_;
    long long iters_ = 0, pops_ = 0, edges_ = 0, relocs_ = 0, pool_used_ = 0;
};

}  // namespace dppr_host
