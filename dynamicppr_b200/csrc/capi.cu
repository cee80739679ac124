// capi.cu -- the extern "C" boundary declared in include/dppr.h.  Exceptions stop here.
#include <cstring>
#include <new>
#include <string>
#include <vector>
// unity build: the kernels live in headers, so the whole library is one translation unit
#include "engine.cu"
#include "streamgen.cuh"

struct dppr_engine {
    dppr::Engine *impl = nullptr;
    std::string err;
};

namespace {
thread_local std::string g_create_error;

template <typename F>
int guarded(dppr_engine *e, F &&f) {
    if (!e || !e->impl) return DPPR_E_INVALID;
    try {
        f(*e->impl);
        return DPPR_OK;
    } catch (const dppr::InvalidArgument &x) {
        e->err = x.what();
        return DPPR_E_INVALID;
    } catch (const dppr::StateError &x) {
        e->err = x.what();
        return DPPR_E_STATE;
    } catch (const dppr::CapacityError &x) {
        e->err = x.what();
        return DPPR_E_CAPACITY;
    } catch (const dppr::CudaFailure &x) {
        e->err = x.what();
        return DPPR_E_CUDA;
    } catch (const std::bad_alloc &) {
        e->err = "host allocation failed";
        return DPPR_E_CAPACITY;
    } catch (const std::exception &x) {
        e->err = x.what();
        return DPPR_E_CUDA;
    }
}
template <typename F>
int guarded_free(F &&f) {
    using namespace dppr;
    try {
        f();
        return DPPR_OK;
    } catch (const InvalidArgument &x) {
        g_create_error = x.what();
        return DPPR_E_INVALID;
    } catch (const CapacityError &x) {
        g_create_error = x.what();
        return DPPR_E_CAPACITY;
    } catch (const std::bad_alloc &) {
        g_create_error = "host allocation failed";
        return DPPR_E_CAPACITY;
    } catch (const std::exception &x) {
        g_create_error = x.what();
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return DPPR_E_NODEVICE;
        return DPPR_E_CUDA;
    }
}
}  // namespace

extern "C" {

int dppr_version(void) { return DPPR_VERSION; }

const char *dppr_last_error(const dppr_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int dppr_create(const dppr_config *cfg, dppr_engine **out) {
    if (!cfg || !out) {
        g_create_error = "null argument";
        return DPPR_E_INVALID;
    }
    *out = nullptr;
    try {
        dppr_engine *h = new dppr_engine();
        try {
            h->impl = new dppr::Engine(*cfg);
        } catch (...) {
            delete h;
            throw;
        }
        *out = h;
        return DPPR_OK;
    } catch (const dppr::InvalidArgument &x) {
        g_create_error = x.what();
        return DPPR_E_INVALID;
    } catch (const dppr::CudaFailure &x) {
        g_create_error = x.what();
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return DPPR_E_NODEVICE;
        return DPPR_E_CUDA;
    } catch (const std::exception &x) {
        g_create_error = x.what();
        return DPPR_E_CUDA;
    }
}

void dppr_destroy(dppr_engine *e) {
    if (!e) return;
    delete e->impl;
    delete e;
}

int dppr_init_window(dppr_engine *e, const int32_t *e1, const int32_t *e2, int64_t n) {
    return guarded(e, [&](dppr::Engine &g) { g.init_window_soa(e1, e2, n); });
}
int dppr_init_window_pairs(dppr_engine *e, const int32_t *pairs, int64_t n) {
    return guarded(e, [&](dppr::Engine &g) { g.init_window_pairs(pairs, n); });
}
int dppr_init_window_device_pairs(dppr_engine *e, const int32_t *dpairs, int64_t n) {
    return guarded(e, [&](dppr::Engine &g) { g.init_window_device_pairs(dpairs, n); });
}
int dppr_solve_initial(dppr_engine *e) {
    return guarded(e, [&](dppr::Engine &g) { g.solve_initial(); });
}
int dppr_apply_batch(dppr_engine *e, const int32_t *e1, const int32_t *e2, int64_t B) {
    return guarded(e, [&](dppr::Engine &g) { g.apply_batch_host_soa(e1, e2, B); });
}
int dppr_apply_batch_pairs(dppr_engine *e, const int32_t *pairs, int64_t B) {
    return guarded(e, [&](dppr::Engine &g) { g.apply_batch_host_pairs(pairs, B); });
}
int dppr_apply_batch_device_pairs(dppr_engine *e, const int32_t *dpairs, int64_t B) {
    return guarded(e, [&](dppr::Engine &g) { g.apply_batch_device_pairs(dpairs, B); });
}
int dppr_refresh(dppr_engine *e) {
    return guarded(e, [&](dppr::Engine &g) { g.refresh(false); });
}
int dppr_slide(dppr_engine *e, const int32_t *e1, const int32_t *e2, int64_t B) {
    return guarded(e, [&](dppr::Engine &g) { g.apply_batch_host_soa(e1, e2, B); g.refresh(false); });
}
int dppr_slide_pairs(dppr_engine *e, const int32_t *pairs, int64_t B) {
    return guarded(e, [&](dppr::Engine &g) { g.apply_batch_host_pairs(pairs, B); g.refresh(false); });
}
int dppr_slide_device_pairs(dppr_engine *e, const int32_t *dpairs, int64_t B) {
    return guarded(e, [&](dppr::Engine &g) { g.apply_batch_device_pairs(dpairs, B); g.refresh(false); });
}
int dppr_sync(dppr_engine *e) {
    return guarded(e, [&](dppr::Engine &g) { g.sync(); });
}
int dppr_wait_event(dppr_engine *e, void *cuda_event) {
    return guarded(e, [&](dppr::Engine &g) { g.wait_event(cuda_event); });
}
int dppr_get_topk(dppr_engine *e, int32_t first, int32_t n, int32_t k, int32_t *ids, double *values) {
    return guarded(e, [&](dppr::Engine &g) { g.topk(first, n, k, ids, values); });
}
int dppr_validate(dppr_engine *e, int32_t s, double *max_abs_residual, double *max_invariant_defect) {
    return guarded(e, [&](dppr::Engine &g) { g.validate(s, max_abs_residual, max_invariant_defect); });
}
int dppr_check_window_device(dppr_engine *e, const int32_t *dpairs, int64_t n, int64_t *mismatches) {
    return guarded(e, [&](dppr::Engine &g) { g.check_window_device(dpairs, n, mismatches); });
}
int dppr_check_window(dppr_engine *e, const int32_t *pairs, int64_t n, int64_t *mismatches) {
    return guarded(e, [&](dppr::Engine &g) {
        if (!pairs || n <= 0) throw dppr::InvalidArgument("null or empty window");
        dppr::DevBuf<int2> tmp;
        tmp.alloc((size_t)n);
        DPPR_CUDA(cudaMemcpy(tmp.ptr, pairs, sizeof(int2) * (size_t)n, cudaMemcpyHostToDevice));
        DPPR_CUDA(cudaDeviceSynchronize());
        g.check_window_device((const int32_t *)tmp.ptr, n, mismatches);
    });
}
int dppr_get_batch_stats(dppr_engine *e, int64_t k, dppr_batch_stats *out) {
    return guarded(e, [&](dppr::Engine &g) { g.get_stats(k, out); });
}
int64_t dppr_batches_done(const dppr_engine *e) { return (e && e->impl) ? e->impl->batches_done() : -1; }
int dppr_get_estimates(dppr_engine *e, int32_t s, double *out) {
    return guarded(e, [&](dppr::Engine &g) { g.get_vector(0, s, out); });
}
int dppr_get_residuals(dppr_engine *e, int32_t s, double *out) {
    return guarded(e, [&](dppr::Engine &g) { g.get_vector(1, s, out); });
}
int dppr_copy_estimates_device(dppr_engine *e, int32_t s, void *dptr) {
    return guarded(e, [&](dppr::Engine &g) { g.copy_estimates_device(s, dptr); });
}
int dppr_export_window_csr(dppr_engine *e, int32_t *rp, int32_t *ci, int32_t *od) {
    return guarded(e, [&](dppr::Engine &g) { g.export_csr(rp, ci, od); });
}
int dppr_export_window_out_csr(dppr_engine *e, int32_t *rp, int32_t *ci) {
    bool have = false;
    const int rc = guarded(e, [&](dppr::Engine &g) {
        have = g.has_out_lists();
        if (have) g.export_csr(rp, ci, nullptr, true);
    });
    return rc != 0 ? rc : (have ? 0 : 1);
}
int64_t dppr_window_csr_entries(const dppr_engine *e) { return (e && e->impl) ? e->impl->csr_entries() : -1; }
int dppr_set_state(dppr_engine *e, int32_t s, const double *p, const double *r) {
    return guarded(e, [&](dppr::Engine &g) { g.set_state(s, p, r); });
}
int dppr_repair_only(dppr_engine *e) {
    return guarded(e, [&](dppr::Engine &g) { g.refresh(true); });
}

int dppr_debug_iterlog(dppr_engine *e, uint32_t *out, int32_t cap, int32_t *n_out) {
    return guarded(e, [&](dppr::Engine &g) { *n_out = g.get_iterlog(out, cap); });
}

int dppr_debug_ctalog(dppr_engine *e, unsigned long long *out, int32_t cap_rows, int32_t *n_out) {
    return guarded(e, [&](dppr::Engine &g) { *n_out = g.get_ctalog(out, cap_rows); });
}

unsigned long long dppr_kernel_launches(void) { return dppr::launch_counter(); }

// ---- synthetic streams: device generator + bit-identical host twin (csrc/streamgen.cuh) --------
int dppr_generate_stream_device(int32_t device, int32_t kind, int32_t V, int64_t first_edge, int64_t n_edges, uint64_t seed,
                                int32_t *device_pairs) {
    using namespace dppr;
    return guarded_free([&]() {
        if (first_edge < 0 || n_edges < 0 || (!device_pairs && n_edges)) throw InvalidArgument("bad arguments");
        const GenParams g = make_gen_params(kind, V, seed);
        DPPR_CUDA(cudaSetDevice(device));
        if (n_edges) {
            gen_stream_kernel<<<148 * 8, kThreads>>>((int2 *)device_pairs, g, first_edge, n_edges); ++launch_counter();
        }
        DPPR_CUDA(cudaGetLastError());
        DPPR_CUDA(cudaDeviceSynchronize());
    });
}

int dppr_generate_stream_host(int32_t kind, int32_t V, int64_t first_edge, int64_t n_edges, uint64_t seed, int32_t *pairs,
                              int32_t threads) {
    using namespace dppr;
    return guarded_free([&]() {
        if (first_edge < 0 || n_edges < 0 || (!pairs && n_edges)) throw InvalidArgument("bad arguments");
        const GenParams g = make_gen_params(kind, V, seed);
        gen_stream_host(g, first_edge, n_edges, (int2 *)pairs, threads);
    });
}

int dppr_generate_rmat_device(int32_t device, int32_t V, int64_t M, uint64_t seed, int32_t *device_pairs) {
    return dppr_generate_stream_device(device, DPPR_STREAM_RMAT, V, 0, M, seed, device_pairs);
}

// ---- source selection: exact degree ranking of a whole stream on the device -----------------
int dppr_rank_by_degree(int32_t device, int32_t V, int32_t directed, int32_t by_out_degree, const int32_t *pairs, int64_t n,
                        int32_t pairs_on_device, int32_t *order, int32_t *out_degree, int32_t *in_degree) {
    using namespace dppr;
    return guarded_free([&]() {
        if (V <= 0 || n < 0 || (!pairs && n) || n > 0x7fffffffll) throw InvalidArgument("bad arguments");
        DPPR_CUDA(cudaSetDevice(device));
        DevBuf<unsigned> od, id;
        DevBuf<int> err;
        od.alloc((size_t)V); id.alloc((size_t)V); err.alloc(1);
        DPPR_CUDA(cudaMemset(od.ptr, 0, od.bytes()));
        DPPR_CUDA(cudaMemset(id.ptr, 0, id.bytes()));
        DPPR_CUDA(cudaMemset(err.ptr, 0, sizeof(int)));
        if (pairs_on_device) {
            if (n) { degree_hist_kernel<<<148 * 8, kThreads>>>((const int2 *)pairs, n, V, directed, od.ptr, id.ptr, err.ptr); ++launch_counter(); }
        } else {
            const int64_t chunk = 1ll << 25;  // 256 MiB of pairs per upload
            DevBuf<int2> buf;
            buf.alloc((size_t)std::min<int64_t>(std::max<int64_t>(n, 1), chunk));
            for (int64_t lo = 0; lo < n; lo += chunk) {
                const int64_t m = std::min(chunk, n - lo);
                DPPR_CUDA(cudaMemcpy(buf.ptr, pairs + 2 * lo, sizeof(int2) * (size_t)m, cudaMemcpyHostToDevice));
                degree_hist_kernel<<<148 * 8, kThreads>>>(buf.ptr, m, V, directed, od.ptr, id.ptr, err.ptr); ++launch_counter();
            }
        }
        DPPR_CUDA(cudaGetLastError());
        int herr = 0;
        DPPR_CUDA(cudaMemcpy(&herr, err.ptr, sizeof(int), cudaMemcpyDeviceToHost));
        if (herr) throw InvalidArgument("edge endpoint outside [0, vertex_count) (workload/Graph.h:101-102 asserts the same)");
        if (order) {
            // descending degree, ties by ascending id (the reference's std::sort leaves tie order unspecified)
            DevBuf<uint32_t> k[2], v[2], scratch;
            for (int i = 0; i < 2; ++i) { k[i].alloc((size_t)V); v[i].alloc((size_t)V); }
            scratch.alloc(sort_scratch_elems(V));
            const int dbits = bits_for((uint64_t)std::max<int64_t>(2 * n, 1));
            const uint32_t degmax = dbits >= 32 ? 0xffffffffu : (uint32_t)((1ull << dbits) - 1);
            const int grid = std::max(1, std::min(div_up(V, kThreads), 148 * 16));
            relabel_keys<<<grid, kThreads>>>(by_out_degree ? od.ptr : id.ptr, degmax, k[0].ptr, v[0].ptr, V); ++launch_counter();
            const int res = sort_pairs(k[0].ptr, v[0].ptr, k[1].ptr, v[1].ptr, V, std::min(dbits, 32), scratch.ptr, 0);
            DPPR_CUDA(cudaGetLastError());
            DPPR_CUDA(cudaDeviceSynchronize());
            DPPR_CUDA(cudaMemcpy(order, v[res].ptr, sizeof(int32_t) * (size_t)V, cudaMemcpyDeviceToHost));
        }
        DPPR_CUDA(cudaDeviceSynchronize());
        if (out_degree) DPPR_CUDA(cudaMemcpy(out_degree, od.ptr, sizeof(int32_t) * (size_t)V, cudaMemcpyDeviceToHost));
        if (in_degree) DPPR_CUDA(cudaMemcpy(in_degree, id.ptr, sizeof(int32_t) * (size_t)V, cudaMemcpyDeviceToHost));
    });
}

// ---- primitive test hooks ------------------------------------------------------------------------
int dppr_test_sort_pairs(int32_t device, uint32_t *keys, uint32_t *vals, int64_t n, int32_t key_bits) {
    using namespace dppr;
    try {
        if (n < 0 || (!keys && n) || (!vals && n)) throw InvalidArgument("bad arguments");
        DPPR_CUDA(cudaSetDevice(device));
        DevBuf<uint32_t> k[2], v[2], scratch;
        for (int i = 0; i < 2; ++i) { k[i].alloc((size_t)n + 1); v[i].alloc((size_t)n + 1); }
        scratch.alloc(sort_scratch_elems(n));
        DPPR_CUDA(cudaMemcpy(k[0].ptr, keys, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
        DPPR_CUDA(cudaMemcpy(v[0].ptr, vals, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
        const int res = sort_pairs(k[0].ptr, v[0].ptr, k[1].ptr, v[1].ptr, n, key_bits, scratch.ptr, 0);
        DPPR_CUDA(cudaGetLastError());
        DPPR_CUDA(cudaDeviceSynchronize());
        DPPR_CUDA(cudaMemcpy(keys, k[res].ptr, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
        DPPR_CUDA(cudaMemcpy(vals, v[res].ptr, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
        return DPPR_OK;
    } catch (const InvalidArgument &x) {
        g_create_error = x.what();
        return DPPR_E_INVALID;
    } catch (const std::exception &x) {
        g_create_error = x.what();
        return DPPR_E_CUDA;
    }
}

uint32_t dppr_test_relabel_slot(uint32_t rank, uint32_t vertex_count, uint32_t blocks) {
    return dppr::relabel_slot(rank, vertex_count, blocks);
}

int dppr_test_exclusive_scan(int32_t device, uint32_t *data, int64_t n, uint64_t *total) {
    using namespace dppr;
    try {
        if (n < 0 || (!data && n)) throw InvalidArgument("bad arguments");
        DPPR_CUDA(cudaSetDevice(device));
        DevBuf<uint32_t> d, scratch, tot;
        d.alloc((size_t)n + 1);
        scratch.alloc(scan_scratch_elems(n));
        tot.alloc(1);
        DPPR_CUDA(cudaMemcpy(d.ptr, data, sizeof(uint32_t) * (size_t)n, cudaMemcpyHostToDevice));
        exclusive_scan<uint32_t>(d.ptr, d.ptr, n, scratch.ptr, tot.ptr, 0);
        DPPR_CUDA(cudaGetLastError());
        DPPR_CUDA(cudaDeviceSynchronize());
        DPPR_CUDA(cudaMemcpy(data, d.ptr, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost));
        uint32_t t = 0;
        DPPR_CUDA(cudaMemcpy(&t, tot.ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (total) *total = t;
        return DPPR_OK;
    } catch (const InvalidArgument &x) {
        g_create_error = x.what();
        return DPPR_E_INVALID;
    } catch (const std::exception &x) {
        g_create_error = x.what();
        return DPPR_E_CUDA;
    }
}

}  // extern "C"
