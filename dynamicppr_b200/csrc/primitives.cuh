// primitives.cuh -- device-wide exclusive scan and stable LSD radix sort of (key, value) pairs.
//
// Hand-written replacements for what the reference takes from Thrust / CUB / cuSPARSE
// (thrust::sort + cusparseXcoo2csr in gpu/SlidingGraphBuilder.cuh:203-221, thrust::exclusive_scan
// :199-200, cub::BlockScan throughout gpu/ExpandRev.cuh).  Both are bandwidth-trivial next to the
// push; they exist so that every per-batch structure is built deterministically (stable order)
// without a single floating-point or order-dependent atomic.
#pragma once
#include "common.cuh"

namespace dppr {

// ---------------------------------------------------------------------------------------------
// exclusive scan, three kernels per level: tile sums -> (recursive) scan of sums -> rescan tiles
// ---------------------------------------------------------------------------------------------
constexpr int kScanItems = 8;
constexpr int kScanTile = kThreads * kScanItems;  // 2048

template <typename T>
__global__ void __launch_bounds__(kThreads) scan_tile_sums(const T *__restrict__ in, T *__restrict__ sums, int64_t n) {
    __shared__ T sm[kWarps + 1];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T acc = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) acc += in[base + k];
    T total;
    block_exclusive_sum(acc, sm, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// rescans each tile and adds the scanned tile offset; when `offsets` is null the launch is a
// single tile.  Writes the grand total (if requested) from the last tile.
template <typename T>
__global__ void __launch_bounds__(kThreads)
    scan_tiles(const T *__restrict__ in, T *__restrict__ out, const T *__restrict__ offsets, int64_t n, T *total_out) {
    __shared__ T sm[kWarps + 1];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T v[kScanItems];
    T acc = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : T(0);
        acc += v[k];
    }
    T total;
    T run = block_exclusive_sum(acc, sm, total) + (offsets ? offsets[blockIdx.x] : T(0));
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == kThreads - 1) *total_out = run;
}

// scratch needed (elements of T) for a scan of n elements
inline size_t scan_scratch_elems(int64_t n) {
    size_t tot = 0;
    while (n > kScanTile) {
        n = div_up(n, kScanTile);
        tot += (size_t)n;
    }
    return tot + 1;
}

// in-place allowed (in == out).  total_out (device pointer) may be null.
template <typename T>
void exclusive_scan(const T *in, T *out, int64_t n, T *scratch, T *total_out, cudaStream_t st) {
    if (n <= 0) {
        if (total_out) DPPR_CUDA(cudaMemsetAsync(total_out, 0, sizeof(T), st));
        return;
    }
    if (n <= kScanTile) {
        scan_tiles<T><<<1, kThreads, 0, st>>>(in, out, nullptr, n, total_out); ++launch_counter();
        return;
    }
    const int tiles = div_up(n, kScanTile);
    scan_tile_sums<T><<<tiles, kThreads, 0, st>>>(in, scratch, n); ++launch_counter();
    exclusive_scan<T>(scratch, scratch, tiles, scratch + tiles, nullptr, st);
    scan_tiles<T><<<tiles, kThreads, 0, st>>>(in, out, scratch, n, total_out); ++launch_counter();
}

// ---------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass
// ---------------------------------------------------------------------------------------------
constexpr int kSortChunks = 8;                              // 32-key chunks per warp per tile
constexpr int kSortTile = kThreads * kSortChunks;           // 2048 keys per CTA
constexpr int kRadix = 256;

// per-tile digit histogram, laid out digit-major: hist[d * tiles + tile]
// (no __restrict__ on the *_tile helpers: the cooperative window kernel reads what other CTAs wrote earlier in
// the same launch, so these loads must stay ordinary coherent loads)
__device__ __forceinline__ void radix_hist_tile(const uint32_t *keys, uint32_t *hist, int64_t n, int shift, int tiles,
                                                int tile, uint32_t *sh /* kRadix */) {
    sh[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)tile * kSortTile;
#pragma unroll
    for (int k = 0; k < kSortChunks; ++k) {
        int64_t i = base + k * kThreads + threadIdx.x;
        if (i < n) atomicAdd(&sh[(keys[i] >> shift) & (kRadix - 1)], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * tiles + tile] = sh[threadIdx.x];
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
    radix_hist(const uint32_t *__restrict__ keys, uint32_t *__restrict__ hist, int64_t n, int shift, int tiles) {
    __shared__ uint32_t sh[kRadix];
    radix_hist_tile(keys, hist, n, shift, tiles, blockIdx.x, sh);
}

// Stable scatter.  Each warp owns a contiguous run of kSortChunks*32 keys and walks it in order;
// __match_any_sync gives every lane its rank among equal digits of the chunk, per-warp running
// digit counts give the rank within the warp's run, a prefix over warps gives the rank within the
// tile, and the scanned histogram gives the tile's base for that digit.
struct RadixScatterSmem {
    uint32_t wcnt[kWarps][kRadix];
    uint32_t gbase[kRadix];
};

__device__ __forceinline__ void radix_scatter_tile(const uint32_t *kin, const uint32_t *vin, uint32_t *kout, uint32_t *vout,
                                                   const uint32_t *hist_scanned, int64_t n, int shift, int tiles, int tile,
                                                   RadixScatterSmem &sm) {
    const unsigned w = warp_id(), l = lane_id();
    for (int i = threadIdx.x; i < kWarps * kRadix; i += kThreads) (&sm.wcnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)tile * kSortTile + (int64_t)w * (kSortChunks * 32);
    uint32_t key[kSortChunks], val[kSortChunks], rnk[kSortChunks];
#pragma unroll
    for (int c = 0; c < kSortChunks; ++c) {
        const int64_t i = wbase + c * 32 + l;
        const bool valid = i < n;
        key[c] = valid ? kin[i] : 0u;
        val[c] = valid ? vin[i] : 0u;
        const uint32_t d = valid ? ((key[c] >> shift) & (kRadix - 1)) : (uint32_t)kRadix;  // invalid lanes group apart
        const unsigned peers = __match_any_sync(kFull, d);
        const uint32_t before = valid ? sm.wcnt[w][d] : 0u;
        __syncwarp();
        rnk[c] = before + __popc(peers & lanemask_lt());
        if (valid && l == (unsigned)(__ffs(peers) - 1)) sm.wcnt[w][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive prefix over warps for digit == threadIdx.x
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < kWarps; ++ww) {
            uint32_t c = sm.wcnt[ww][d];
            sm.wcnt[ww][d] = run;
            run += c;
        }
        sm.gbase[d] = hist_scanned[(size_t)d * tiles + tile];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kSortChunks; ++c) {
        const int64_t i = wbase + c * 32 + l;
        if (i < n) {
            const uint32_t d = (key[c] >> shift) & (kRadix - 1);
            const uint32_t pos = sm.gbase[d] + sm.wcnt[w][d] + rnk[c];
            kout[pos] = key[c];
            vout[pos] = val[c];
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
    radix_scatter(const uint32_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint32_t *__restrict__ kout,
                  uint32_t *__restrict__ vout, const uint32_t *__restrict__ hist_scanned, int64_t n, int shift,
                  int tiles) {
    __shared__ RadixScatterSmem sm;
    radix_scatter_tile(kin, vin, kout, vout, hist_scanned, n, shift, tiles, blockIdx.x, sm);
}

inline size_t sort_scratch_elems(int64_t n) {
    const size_t h = (size_t)kRadix * (size_t)div_up(n > 0 ? n : 1, kSortTile);
    return h + scan_scratch_elems((int64_t)h);
}

// Sorts n pairs by the low `key_bits` bits of the key, stably.  (k0, v0) hold the input; (k1, v1)
// are same-sized alternates.  Returns 0 if the result is in (k0, v0), 1 if in (k1, v1).
// n must be < 2^32 (positions are 32-bit).  scratch: sort_scratch_elems(n) uint32.
inline int sort_pairs(uint32_t *k0, uint32_t *v0, uint32_t *k1, uint32_t *v1, int64_t n, int key_bits,
                      uint32_t *scratch, cudaStream_t st) {
    if (n <= 1 || key_bits <= 0) return 0;
    const int tiles = div_up(n, kSortTile);
    const int passes = (key_bits + 7) / 8;
    uint32_t *hist = scratch;
    uint32_t *scan_scratch = scratch + (size_t)kRadix * tiles;
    uint32_t *kin = k0, *vin = v0, *kout = k1, *vout = v1;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        radix_hist<<<tiles, kThreads, 0, st>>>(kin, hist, n, shift, tiles); ++launch_counter();
        exclusive_scan<uint32_t>(hist, hist, (int64_t)kRadix * tiles, scan_scratch, nullptr, st);
        radix_scatter<<<tiles, kThreads, 0, st>>>(kin, vin, kout, vout, hist, n, shift, tiles); ++launch_counter();
        uint32_t *t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    return passes & 1;
}

inline int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b) != 0) ++b;
    return b;
}

}  // namespace dppr
