// capi.cu -- the extern "C" boundary declared in include/dppr.h.  Exceptions stop here.
#include <cstring>
#include <new>
#include <string>
#include <vector>
// unity build: the kernels live in headers, so the whole library is one translation unit
#include "engine.cu"

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

// ---- synthetic stream generator on the device -------------------------------------------
namespace {
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
// R-MAT (a, b, c, d) with `scale` levels, counter-based: edge i is a pure function of (seed, i)
__global__ void rmat_kernel(int2 *out, long long M, int V, int scale, unsigned long long seed, unsigned ta, unsigned tab,
                            unsigned tabc, unsigned long long mult, unsigned long long add) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long src = 0, dst = 0;
        unsigned long long h = 0;
        for (int l = 0; l < scale; ++l) {
            if ((l & 1) == 0) h = mix64(seed ^ ((unsigned long long)i * 0x100000001b3ull + (unsigned long long)(l >> 1)));
            const unsigned r = (l & 1) ? (unsigned)(h >> 32) : (unsigned)h;  // 32 uniform bits per level
            const unsigned sbit = r >= tab, dbit = (r >= ta && r < tab) || r >= tabc;
            src = (src << 1) | sbit;
            dst = (dst << 1) | dbit;
        }
        src %= (unsigned long long)V; dst %= (unsigned long long)V;
        out[i] = make_int2((int)((src * mult + add) % (unsigned long long)V), (int)((dst * mult + add) % (unsigned long long)V));
    }
}
}  // namespace

int dppr_generate_rmat_device(int32_t device, int32_t V, int64_t M, uint64_t seed, int32_t *device_pairs) {
    using namespace dppr;
    try {
        if (V <= 1 || M < 0 || !device_pairs) throw InvalidArgument("bad arguments");
        DPPR_CUDA(cudaSetDevice(device));
        int scale = 1;
        while ((1ll << scale) < (long long)V) ++scale;
        unsigned long long mult = 2654435761ull % (unsigned long long)V;
        auto gcd = [](unsigned long long a, unsigned long long b) { while (b) { unsigned long long t = a % b; a = b; b = t; } return a; };
        while (gcd(mult, (unsigned long long)V) != 1) ++mult;
        const double a = 0.57, b = 0.19, c = 0.19;
        const unsigned ta = (unsigned)(a * 4294967296.0), tab = (unsigned)((a + b) * 4294967296.0),
                       tabc = (unsigned)((a + b + c) * 4294967296.0);
        rmat_kernel<<<148 * 8, 256>>>((int2 *)device_pairs, M, V, scale, seed, ta, tab, tabc, mult, mix64(seed) % (unsigned long long)V);
        DPPR_CUDA(cudaGetLastError());
        DPPR_CUDA(cudaDeviceSynchronize());
        return DPPR_OK;
    } catch (const InvalidArgument &x) {
        g_create_error = x.what();
        return DPPR_E_INVALID;
    } catch (const std::exception &x) {
        g_create_error = x.what();
        return DPPR_E_CUDA;
    }
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
