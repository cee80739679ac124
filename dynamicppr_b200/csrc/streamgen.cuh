// streamgen.cuh -- seeded synthetic edge streams, generated on the device AND by a bit-identical host twin
// (stands in for the reference's offline encoder, encoder/GraphEncoder.h:20-98, for shapes
// that are too large to ship: the Twitter-shaped stream is 11.7 GB).
//
// Counter-based: edge i is a pure function of (kind, V, seed, i), computed in INTEGER arithmetic only, so
//   * any slice [first, first + n) can be produced independently (the GPU generates the whole stream in HBM, the
//     host writes just the prefix the reference CPU implementation will read into a sparse .bin file),
//   * the device kernel and the host loop below produce the same bytes (tests/test_gpu_streamgen.py).
// The few floating-point constants (R-MAT quadrant thresholds, the power-law lower cut) are evaluated once, on
// the host, by make_gen_params() for both paths.
//
// Shapes:
//   kRmat      directed R-MAT, (a, b, c, d) = (0.57, 0.19, 0.19, 0.05), ceil(log2 V) levels, ids folded to < V and
//              scrambled by an affine bijection; duplicates and self-loops kept (the reference keeps multi-edges).
//   kPowerLaw  undirected power-law (Chung-Lu endpoint sampling): both endpoints drawn independently with
//              P(rank k) ~ (k + 1)^-0.75 by inverting the continuous CDF in fixed point,
//              rank = floor((V + 1) t^4) - 1, t uniform on [t0, 1), t0 = (V + 1)^-1/4; ranks scrambled by the same
//              kind of bijection; self-loops re-drawn; duplicate pairs are possible and kept.
#pragma once
#include <thread>
#include <vector>
#include "common.cuh"

namespace dppr {

enum : int { kGenRmat = 0, kGenPowerLaw = 1 };

struct GenParams {
    int kind;
    int V;
    int scale;                          // R-MAT levels
    unsigned long long seed;
    unsigned ta, tab, tabc;             // R-MAT quadrant thresholds in 0.32 fixed point
    unsigned long long mult, add;       // id scramble: v -> (v * mult + add) mod V, gcd(mult, V) = 1
    unsigned long long t0, one_minus_t0;  // power law: lower cut of t in 0.32 fixed point, and 2^32 - t0
};

__host__ __device__ __forceinline__ unsigned long long gen_mix64(unsigned long long x) {  // splitmix64 finaliser
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

__host__ __device__ __forceinline__ unsigned long long gen_mulhi64(unsigned long long a, unsigned long long b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (unsigned long long)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// rank of one power-law endpoint from 32 uniform bits
__host__ __device__ __forceinline__ unsigned long long gen_powerlaw_rank(const GenParams &g, unsigned u) {
    const unsigned long long t = g.t0 + (((unsigned long long)u * g.one_minus_t0) >> 32);  // 0.32 fixed point, < 2^32
    const unsigned long long t2 = t * t;                       // 0.64
    const unsigned long long t4 = gen_mulhi64(t2, t2);         // 0.64
    const unsigned long long x = gen_mulhi64(t4, (unsigned long long)g.V + 1ull);  // floor((V + 1) t^4)
    unsigned long long k = x >= 1ull ? x - 1ull : 0ull;
    if (k >= (unsigned long long)g.V) k = (unsigned long long)g.V - 1ull;
    return k;
}

__host__ __device__ __forceinline__ int2 gen_edge(const GenParams &g, unsigned long long i) {
    const unsigned long long V = (unsigned long long)g.V;
    if (g.kind == kGenRmat) {
        unsigned long long src = 0, dst = 0, h = 0;
        for (int l = 0; l < g.scale; ++l) {
            if ((l & 1) == 0) h = gen_mix64(g.seed ^ (i * 0x100000001b3ull + (unsigned long long)(l >> 1)));
            const unsigned r = (l & 1) ? (unsigned)(h >> 32) : (unsigned)h;  // 32 uniform bits per level
            const unsigned sbit = r >= g.tab, dbit = (r >= g.ta && r < g.tab) || r >= g.tabc;
            src = (src << 1) | sbit;
            dst = (dst << 1) | dbit;
        }
        src %= V; dst %= V;
        return make_int2((int)((src * g.mult + g.add) % V), (int)((dst * g.mult + g.add) % V));
    }
    unsigned long long a = 0, b = 0;
    for (unsigned attempt = 0; attempt < 4u; ++attempt) {
        const unsigned long long h = gen_mix64(g.seed ^ (i * 0x100000001b3ull + 0x51ed270b0000ull + attempt));
        a = gen_powerlaw_rank(g, (unsigned)h);
        b = gen_powerlaw_rank(g, (unsigned)(h >> 32));
        if (a != b) break;
    }
    if (a == b) b = (a + 1ull) % V;  // (never a self-loop)
    return make_int2((int)((a * g.mult + g.add) % V), (int)((b * g.mult + g.add) % V));
}

inline GenParams make_gen_params(int kind, int V, unsigned long long seed) {
    if (kind != kGenRmat && kind != kGenPowerLaw) throw InvalidArgument("stream kind must be DPPR_STREAM_RMAT or DPPR_STREAM_POWERLAW");
    if (V <= 1) throw InvalidArgument("vertex_count must be at least 2");
    GenParams g{};
    g.kind = kind;
    g.V = V;
    g.seed = seed;
    g.scale = 1;
    while ((1ll << g.scale) < (long long)V) ++g.scale;
    const double a = 0.57, b = 0.19, c = 0.19;
    g.ta = (unsigned)(a * 4294967296.0);
    g.tab = (unsigned)((a + b) * 4294967296.0);
    g.tabc = (unsigned)((a + b + c) * 4294967296.0);
    auto gcd = [](unsigned long long x, unsigned long long y) { while (y) { const unsigned long long t = x % y; x = y; y = t; } return x; };
    g.mult = 2654435761ull % (unsigned long long)V;
    if (g.mult == 0) g.mult = 1;
    while (gcd(g.mult, (unsigned long long)V) != 1) ++g.mult;
    g.add = gen_mix64(seed) % (unsigned long long)V;
    // t0 = (V + 1)^-1/4 by integer bisection on t^4 (V + 1) >= 1 in the generator's own fixed point (no libm, so
    // every build of this file agrees on it)
    unsigned long long lo = 0, hi = 0xffffffffull;
    while (hi - lo > 1) {
        const unsigned long long mid = (lo + hi) >> 1, m2 = mid * mid;
        if (gen_mulhi64(gen_mulhi64(m2, m2), (unsigned long long)V + 1ull) >= 1ull) hi = mid; else lo = mid;
    }
    g.t0 = hi;
    g.one_minus_t0 = 0x100000000ull - g.t0;
    return g;
}

__global__ void __launch_bounds__(kThreads) gen_stream_kernel(int2 *__restrict__ out, GenParams g, long long first, long long n) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads)
        out[i] = gen_edge(g, (unsigned long long)(first + i));
}

inline void gen_stream_host(const GenParams &g, long long first, long long n, int2 *out, int threads) {
    threads = threads < 1 ? 1 : (threads > 256 ? 256 : threads);
    if (n < (1 << 16)) threads = 1;
    std::vector<std::thread> pool;
    const long long per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        const long long lo = (long long)t * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) break;
        pool.emplace_back([&g, first, out, lo, hi]() {
            for (long long i = lo; i < hi; ++i) out[i] = gen_edge(g, (unsigned long long)(first + i));
        });
    }
    for (auto &th : pool) th.join();
}

// ---- degree ranking (reference: workload/Graph.h:85-131 histograms + :178-198 sort by degree) ---------------------
__global__ void __launch_bounds__(kThreads)
    degree_hist_kernel(const int2 *__restrict__ pairs, long long n, int V, int directed, unsigned *__restrict__ outdeg,
                       unsigned *__restrict__ indeg, int *errflag) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
        const int2 e = pairs[i];
        if ((unsigned)e.x >= (unsigned)V || (unsigned)e.y >= (unsigned)V) { atomicOr(errflag, 1); continue; }
        atomicAdd(&outdeg[e.x], 1u);
        atomicAdd(&indeg[e.y], 1u);
        if (!directed) { atomicAdd(&outdeg[e.y], 1u); atomicAdd(&indeg[e.x], 1u); }
    }
}

}  // namespace dppr
