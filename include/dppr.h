/* dppr.h -- C ABI of the B200-native streaming reverse-push PPR engine.
 *
 * This is the drop-in boundary for the hot path of guowentian/dynamicppr.
 * The reference has no FFI layer; what `gpu/PPRGPUMain.cu` calls is a C++ class surface
 * (PPRGPU / PPRRevPushGPU{,FF,Eager,Vanilla}, SlidingGraphBuilder, DeviceMemory).  Every entry
 * point below names the reference interface it replaces (file:line relative to the reference
 * root).  Plain C types only: pointers, sizes, an opaque handle.  No exceptions, no exit() and
 * no torch types cross this boundary; every function returns 0 on success or a DPPR_E_* code,
 * with text available from dppr_last_error().
 *
 * Ownership: the engine owns every device allocation (reference: DeviceMemory owns all arrays,
 * gpu/DeviceMemory.cuh:9-135).  Host pointers passed in are borrowed for the duration of the
 * call only (they are staged into pinned memory before the call returns).
 * Threading: one handle = one GPU = one CUDA stream; a handle is not thread-safe; distinct
 * handles may be driven from distinct host threads or processes (one process per GPU is how
 * bench.py shards sources).
 * Asynchrony: dppr_slide*, dppr_apply_batch* and dppr_refresh only enqueue work; call
 * dppr_sync() (or any dppr_get_* / dppr_export_* call, which synchronise) before reading
 * results on the host.
 */
#ifndef DPPR_H
#define DPPR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPPR_VERSION 200

enum {
    DPPR_OK = 0,
    DPPR_E_INVALID = 1,   /* bad argument / configuration (reference: ArgumentsChecker, Arguments.h:42-64) */
    DPPR_E_CUDA = 2,      /* CUDA runtime failure (reference: CUDA_ERROR prints + exit(-1), gpu/GPUUtil.cuh:7-19) */
    DPPR_E_STATE = 3,     /* call out of order (e.g. slide before init_window) */
    DPPR_E_CAPACITY = 4,  /* adjacency pool / frontier queue exhausted -- enlarge via dppr_config.  Detected on the device:
                             reported by the first call after the failing batch has finished (dppr_slide* / dppr_apply_batch* /
                             dppr_sync / dppr_get_*), and by every call after that -- the engine stays failed */
    DPPR_E_NODEVICE = 5   /* no usable CUDA device: there is NO CPU fallback */
};

/* the four push variants of the reference CLI flag -o (Meta.h:11-17) */
enum { DPPR_OPTIMIZED = 0, DPPR_FAST_FRONTIER = 1, DPPR_EAGER = 2, DPPR_VANILLA = 3 };

/* how the push loop is driven */
enum {
    DPPR_ENGINE_AUTO = 0,      /* = LEVELSYNC */
    DPPR_ENGINE_STEPWISE = 1,  /* one launch per push sub-pass, host reads the frontier count every iteration
                                  (the reference's structure, gpu/PPRRevPushGPU.cuh:97-131); debugging / profiling */
    /* 2 was an experimental ticket-queue engine in round 1; removed (it lost on every probe), the value is rejected */
    DPPR_ENGINE_LEVELSYNC = 3  /* one cooperative launch per refresh, level-synchronous iterations separated by a
                                  software grid barrier (csrc/push.cuh) */
};

/* Tuning knobs.  Every field: 0 = default.  Read once by dppr_create; a DPPR_<NAME> environment variable of the same
 * meaning (development aid for A/B runs, also read once, in dppr_create) is consulted only where the field is 0. */
typedef struct dppr_tuning {
    int32_t relabel;            /* internal vertex order by out-degree rank: 0/1 on (default), -1 off             [DPPR_RELABEL=0] */
    int32_t relabel_blocks;     /* ranks are dealt round-robin over this many blocks; default 1024               [DPPR_RELABEL_BLOCKS] */
    int32_t relabel_both;       /* 1: rank by out- plus in-degree                                                 [DPPR_RELABEL_BOTH] */
    int32_t ctas_per_sm;        /* co-resident CTAs per SM of the persistent push kernel; default 4               [DPPR_CTAS_PER_SM] */
    int32_t tile_cap;           /* frontier items per tile once a CTA's share exceeds 512; default 128            [DPPR_TILE_CAP] */
    int32_t max_iters;          /* push iterations per refresh before the watchdog fires; default 400000         [DPPR_MAX_ITERS] */
    int32_t dense;              /* gather sweeps (variant 0): 0 auto by window size, 1 always, -1 never           [DPPR_DENSE] */
    int32_t pull_group;         /* most lanes sharing a vertex in a multi-source sweep (1..16); default 16        [DPPR_PULL_GROUP] */
    int32_t pull_warp_min;      /* out-lists from this length are walked by a whole warp; default 32              [DPPR_PULL_WARP_MIN] */
    int32_t pull_big_min;       /* ... and from this length cut into chunks for the grid; default 4096 / 1024 (1 / several sources) [DPPR_PULL_BIG_MIN] */
    int32_t pull_big_chunk;     /* entries per chunk; default pull_big_min / 4                                    [DPPR_PULL_BIG_CHUNK] */
    int32_t window_path;        /* 0 auto, 1 multi-kernel window update only, 2 no single-CTA kernel              [DPPR_WINDOW_PATH] */
    int32_t iterlog;            /* 1: keep a per-iteration log of the last refresh (dppr_debug_iterlog)           [DPPR_ITERLOG] */
    int32_t probe_iter;         /* iteration whose per-CTA timeline dppr_debug_ctalog returns; default 10         [DPPR_PROBE_ITER] */
    double dense_div;           /* static scatter/gather switch estimate: (E_w + 2V) S / dense_div; default 4     [DPPR_DENSE_DIV] */
    double dense_min_edges;     /* auto: window entries x sources from which the switching kernel is used; default 2e7 for
                                   variant 0, always for variants 1-3                                             [DPPR_DENSE_MIN_EDGES] */
    double carry_gamma;         /* variant 0 threshold schedule; default off (1.0)                               [DPPR_CARRY_GAMMA] */
    double carry_scale;         /* default 0.01                                                                   [DPPR_CARRY_SCALE] */
    int32_t dense_accel;        /* Chebyshev-accelerated sweeps on undirected windows: 0/1 on (default), -1 off       [DPPR_DENSE_ACCEL] */
    int32_t signed_push;        /* variant 0: one pass over both residual signs: 0/1 on (default), -1 the reference's two passes [DPPR_SIGNED_PUSH] */
    int32_t panel_sources;      /* most sources refreshed by one launch; more are split into equal panels, each with its own
                                   contiguous [V][panel] state, refreshed one after the other; default 128            [DPPR_PANEL_SOURCES] */
    int32_t pull_warp_units;    /* multi-source sweeps: work is handed to warps, this many warp-slices (32 / lanes-per-vertex vertices each) at a
                                   time; default 8 = one tile.  -1: to whole CTAs (a barrier per item)                [DPPR_PULL_WARP_UNITS] */
    int32_t reserved[4];
} dppr_tuning;

typedef struct dppr_engine dppr_engine;

typedef struct dppr_config {
    int32_t vertex_count;       /* V: first int32 of the .bin file (SlidingGraphVec.h:42) */
    int32_t directed;           /* -i: 1 directed, 0 undirected (edges mirrored on device) */
    int64_t window_edges;       /* W: stream edges in the window (SlidingGraphVec.h:48) */
    int64_t max_batch_edges;    /* largest B that will be passed to dppr_slide / dppr_apply_batch */
    double alpha;               /* teleport probability; the reference fixes 0.15 (Meta.h:31). <=0 -> 0.15 */
    double epsilon;             /* -e: residual tolerance (Arguments.h:83). <=0 -> 1e-9 */
    int32_t variant;            /* -o: DPPR_OPTIMIZED .. DPPR_VANILLA */
    int32_t device;             /* CUDA device ordinal */
    int32_t n_sources;          /* >=1; one (p, r) pair per source, all sharing the window graph */
    const int32_t *sources;     /* n_sources vertex ids (-s; the reference runs one per process) */
    int32_t engine_mode;        /* DPPR_ENGINE_* */
    int32_t record_timing;      /* 1: bracket every batch phase with CUDA events (see dppr_batch_stats) */
    double pool_factor;         /* adjacency pool slots per window CSR entry; <=0 -> default (8.0) */
    int64_t frontier_capacity;  /* (source, vertex) items per frontier queue; <=0 -> default */
    int32_t hub_degree;         /* in-degree at/above which a vertex is expanded grid-wide; <=0 -> 64 */
    int32_t reserved0;
    dppr_tuning tuning;
} dppr_config;

typedef struct dppr_batch_stats {
    int64_t batch_index;        /* 0 = initial solve, k = k-th slide */
    int64_t edges;              /* B stream edges slid in this batch */
    int64_t batch_entries;      /* N_b = 2*D*B directed insert/delete entries */
    int64_t touched_vertices;   /* distinct u whose residual was repaired (seed candidates) */
    int64_t iterations;         /* push iterations, both phases */
    int64_t frontier_pops;      /* F: (source, vertex) pops */
    int64_t traversed_edges;    /* T: in-edges traversed = FP64 atomics issued */
    int64_t hub_pops;           /* pops expanded grid-wide */
    int64_t relocations;        /* adjacency rings moved to a larger slot range in this batch */
    int64_t pool_used;          /* adjacency pool high-water mark (slots) */
    float ms_upload;            /* H2D of the batch (0 for device-resident input) */
    float ms_window;            /* device window update (sort + expire + insert) */
    float ms_repair;            /* residual repair   } reference "ppr_time" = ms_repair + ms_push */
    float ms_push;              /* both push phases  } (gpu/PPRGPU.cuh:128-163)                    */
    int32_t error_flags;        /* device-side DPPR_DEVERR_* bits, 0 when healthy */
    int32_t dense_sweeps;       /* iterations that ran as gather sweeps over the out-lists (counted in `iterations` too) */
    /* traversed_edges = scatter_edges + the (edge, source) pairs the sweeps gathered with a non-zero residual, i.e. the
     * work in the reference's push form.  What the sweeps actually moved (they read every out-list entry of the active
     * tiles, whatever the frontier) is counted separately, for the roofline of the gather form: */
    int64_t scatter_edges;      /* in-edges traversed by scatter iterations = FP64 atomics issued */
    int64_t dense_slots;        /* out-list entries walked by sweeps (once per chunk group of sources) */
    int64_t dense_pairs;        /* (out-list entry, source) gathers of x by sweeps */
    int64_t dense_units;        /* (vertex, source) units finished by sweeps */
    int64_t dense_pops;         /* frontier pops performed by sweeps (included in frontier_pops) */
    int64_t pool_leaked;        /* pool slots dropped because a free stack was full (0 in every run so far) */
} dppr_batch_stats;

enum {
    DPPR_DEVERR_POOL = 1,       /* adjacency pool exhausted */
    DPPR_DEVERR_QUEUE = 2,      /* frontier queue overflow */
    DPPR_DEVERR_HUBQ = 4,       /* hub list overflow */
    DPPR_DEVERR_WATCHDOG = 8,   /* grid barrier / iteration watchdog fired */
    DPPR_DEVERR_UNDERFLOW = 16, /* expiry of an edge the window does not hold (caller broke FIFO order) */
    DPPR_DEVERR_BADID = 32      /* an edge endpoint outside [0, vertex_count) reached the device (device-pointer input; host
                                   input is checked before it is staged and fails the call with DPPR_E_INVALID) */
};

int dppr_version(void);
/* text of the last failure on this handle (or of the last failed dppr_create when e == NULL) */
const char *dppr_last_error(const dppr_engine *e);

/* Replaces: new PPRRevPushGPU*(graph) -- PPRGPU::PPRGPU (gpu/PPRGPU.cuh:23-34), DeviceMemory
 * allocation (gpu/DeviceMemory.cuh:52-74), new SlidingGraphBuilder (gpu/SlidingGraphBuilder.cuh:64-76)
 * and the variant choice of gpu/PPRGPUMain.cu:24-27. */
int dppr_create(const dppr_config *cfg, dppr_engine **out);
void dppr_destroy(dppr_engine *e);

/* Replaces: SlidingGraphBuilder::InitWindowStream (gpu/SlidingGraphBuilder.cuh:182-192) +
 * DeviceMemory::CudaMemcpyGraph/CudaMemcpyRowPtr (gpu/DeviceMemory.cuh:76-108): loads the first W
 * stream edges and builds the device-resident window graph.  SoA form mirrors EdgeBatch
 * (EdgeBatch.h:6-30); the pairs form takes the .bin payload as it lies in the file. */
int dppr_init_window(dppr_engine *e, const int32_t *edge1, const int32_t *edge2, int64_t n);
int dppr_init_window_pairs(dppr_engine *e, const int32_t *pairs, int64_t n);
/* same, the W edges already resident in device memory (W x int32 pairs) */
int dppr_init_window_device_pairs(dppr_engine *e, const int32_t *device_pairs, int64_t n);

/* Replaces: Init<<<>>> + ExecuteMainLoop(0) of PPRGPU::DynamicExecute (gpu/PPRGPU.cuh:84-89):
 * r[s]=1, p=0, push phase 0 to exhaustion on the initial window, for every source. */
int dppr_solve_initial(dppr_engine *e);

/* Replaces: GPUEdgeBatch::CudaMemcpy (gpu/GPUEdgeBatch.cuh:20-26) + GPUBuildSlidingGraph ->
 * SlidingGraphBuilder::IncBuildInGraph (gpu/PPRRevPushGPU.cuh:38-44, gpu/SlidingGraphBuilder.cuh:117-133).
 * Takes only the B ARRIVING edges (the reference's `new_stream`, SlidingGraphVec.h:226-236); the B
 * expiring edges are read from the device's own arrival-order ring.  O(B) work, no re-sort of the window. */
int dppr_apply_batch(dppr_engine *e, const int32_t *new_edge1, const int32_t *new_edge2, int64_t B);
int dppr_apply_batch_pairs(dppr_engine *e, const int32_t *pairs, int64_t B);
/* same, input already resident in device memory (B x int32 pairs).  Stream contract of every *_device_pairs entry point:
 * the engine reads the buffer asynchronously on its OWN non-blocking stream, which does not order against the stream
 * that produced the buffer (not even the legacy default stream).  The caller must either have synchronised the producer
 * before the call, or pass an event recorded after the producer to dppr_wait_event() first; and it must keep the buffer
 * alive and unmodified until dppr_sync() (or any synchronising call) returns. */
int dppr_apply_batch_device_pairs(dppr_engine *e, const int32_t *device_pairs, int64_t B);
/* makes the engine's stream wait for `cuda_event` (a cudaEvent_t) before any work enqueued after this call */
int dppr_wait_event(dppr_engine *e, void *cuda_event);

/* Replaces: IncrementalBatchUpdate + ExecuteMainLoop(0) + ExecuteMainLoop(1), the region the reference
 * times as ppr_time (gpu/PPRGPU.cuh:128-163; kernels gpu/StreamUpdate.cuh, Inspect.cuh, ExpandRev.cuh). */
int dppr_refresh(dppr_engine *e);

/* apply_batch + refresh: one streaming step (body of SlidingWindowExecuteMainLoop, gpu/PPRGPU.cuh:109-169) */
int dppr_slide(dppr_engine *e, const int32_t *new_edge1, const int32_t *new_edge2, int64_t B);
int dppr_slide_pairs(dppr_engine *e, const int32_t *pairs, int64_t B);
int dppr_slide_device_pairs(dppr_engine *e, const int32_t *device_pairs, int64_t B);

int dppr_sync(dppr_engine *e);
/* kernels launched by this library from the calling thread so far (reported by bench.py as gpu_launches) */
unsigned long long dppr_kernel_launches(void);

/* stats of batch `batch_index` (0 = initial solve); -1 = most recent.  Synchronises. */
int dppr_get_batch_stats(dppr_engine *e, int64_t batch_index, dppr_batch_stats *out);
int64_t dppr_batches_done(const dppr_engine *e);

/* Replaces: the cudaMemcpy D2H of pagerank / residual in ValidateResult (gpu/PPRRevPushGPU.cuh:134-139).
 * out has V doubles.  Synchronises. */
int dppr_get_estimates(dppr_engine *e, int32_t source_index, double *out);
int dppr_get_residuals(dppr_engine *e, int32_t source_index, double *out);
/* device-to-device copy of one estimate vector into caller-owned device memory (for the final
 * NCCL gather done by the caller; there is no collective on the hot path). */
int dppr_copy_estimates_device(dppr_engine *e, int32_t source_index, void *device_out);
/* What a query reads back: the k (<= 128) largest estimates of sources [first_source, first_source + n_sources), selected
 * on the device (csrc/topk.cuh).  ids / values: n_sources x k, host memory; value descending, ties by ascending vertex id;
 * rows shorter than k (V < k) are padded with id -1.  Synchronises. */
int dppr_get_topk(dppr_engine *e, int32_t first_source, int32_t n_sources, int32_t k, int32_t *ids, double *values);

/* ---- validation on the device (reference: -DVALIDATE, gpu/PPRRevPushGPU.cuh:45-90,133-165; csrc/validate.cuh) ---------
 * dppr_validate: max |r[u]| (the reference asserts < eps) and the largest defect of the push invariant
 *   p[u] + a r[u] = a [u==s] + (1-a)/(outdeg(u)+1) sum_{w in out(u)} p[w]   over all u, from the device-resident window
 * graph; defect ~1e-15 and max |r| <= eps together imply |p - pi| <= eps.  Either output may be NULL.  Synchronises.
 * dppr_check_window_device: the reference's ValidateGraph, without the host: `device_pairs` = the n = W edges the window
 * must hold right now (device memory, stream order, caller ids); *mismatches = entries of the canonical (dst, src)-sorted
 * edge list + out-degrees that differ from the engine's window graph (0 = bit-exact). */
int dppr_validate(dppr_engine *e, int32_t source_index, double *max_abs_residual, double *max_invariant_defect);
int dppr_check_window_device(dppr_engine *e, const int32_t *device_pairs, int64_t n, int64_t *mismatches);
/* same, the W window edges in host memory (e.g. the mmap'ed file: records [k*B, k*B + W) after k batches) */
int dppr_check_window(dppr_engine *e, const int32_t *pairs, int64_t n, int64_t *mismatches);

/* Canonical window graph, the object the reference validator compares
 * (gpu/PPRRevPushGPU.cuh:45-90): in_row_ptr[V+1], in_col_ind[E_w] with rows ascending and duplicates
 * kept, out_deg[V] (reference row_ptr differences).  Any pointer may be NULL.  Test/export path:
 * sorts on the device; not part of the hot path. */
int dppr_export_window_csr(dppr_engine *e, int32_t *in_row_ptr, int32_t *in_col_ind, int32_t *out_deg);
int64_t dppr_window_csr_entries(const dppr_engine *e); /* E_w = D*W */
/* Test hook: the transposed window graph the dense (gather) iterations read -- out_row_ptr[V+1], out_col_ind[E_w],
 * rows ascending, duplicates kept.  Returns 1 and fills nothing if this engine keeps no separate out-lists
 * (undirected graph: the in-lists serve; or dense iterations disabled). */
int dppr_export_window_out_csr(dppr_engine *e, int32_t *out_row_ptr, int32_t *out_col_ind);

/* ---- synthetic streams ------------------------------------------------------------------
 * Stand-in for the reference's offline encoder (encoder/GraphEncoder.h:20-98) for shapes too large to ship: the
 * payload of a .bin file (int32 pairs in stream order, encoder/GraphEncoder.h:86-95) generated from a seed.
 * Counter-based and integer-only: edge i depends on (kind, vertex_count, seed, i) alone, so any slice
 * [first_edge, first_edge + n_edges) can be produced on its own, and the DEVICE generator and its HOST twin write
 * identical bytes -- the GPU holds the whole stream in HBM while the host writes only the prefix the reference CPU
 * implementation reads.  DPPR_STREAM_RMAT: directed R-MAT (0.57, 0.19, 0.19, 0.05), ids folded and scrambled into
 * [0, V), duplicates and self-loops kept.  DPPR_STREAM_POWERLAW: undirected Chung-Lu power law, P(rank k) ~
 * (k+1)^-0.75 per endpoint, no self-loops, duplicate pairs kept (the reference keeps multi-edges). */
enum { DPPR_STREAM_RMAT = 0, DPPR_STREAM_POWERLAW = 1 };
int dppr_generate_stream_device(int32_t device, int32_t kind, int32_t vertex_count, int64_t first_edge, int64_t n_edges,
                                uint64_t seed, int32_t *device_pairs);
/* host twin: no GPU needed; `threads` host threads */
int dppr_generate_stream_host(int32_t kind, int32_t vertex_count, int64_t first_edge, int64_t n_edges, uint64_t seed,
                              int32_t *pairs, int32_t threads);
/* = dppr_generate_stream_device(device, DPPR_STREAM_RMAT, vertex_count, 0, n_edges, seed, device_pairs) */
int dppr_generate_rmat_device(int32_t device, int32_t vertex_count, int64_t n_edges, uint64_t seed, int32_t *device_pairs);

/* ---- source selection ---------------------------------------------------------------------
 * Replaces: workload/Graph.h:85-131 (degree histograms over the whole file; an undirected edge counts at both ends)
 * + the sort of ChooseVertexDegreeRange (workload/Graph.h:178-190).  `pairs` = n int32 pairs in host memory
 * (uploaded in chunks) or, with pairs_on_device != 0, in device memory.  order[V] (host): vertex ids by descending
 * out-degree (by_out_degree != 0) or in-degree, ties by ascending id -- exact, also at Twitter scale.  out_degree[V] /
 * in_degree[V] (host, optional): the histograms.  The bucket rules on top of the ranking (top10 = ranks 0..9, top1000 =
 * 10 random ranks of [10, 1000), ...; workload/Workload.cpp:47-55) live in the `pagerank --pick` host tool. */
int dppr_rank_by_degree(int32_t device, int32_t vertex_count, int32_t directed, int32_t by_out_degree, const int32_t *pairs,
                        int64_t n_edges, int32_t pairs_on_device, int32_t *order, int32_t *out_degree, int32_t *in_degree);

/* ---- test hooks (used by tests/ only) ---------------------------------------------------- */
/* overwrite (p, r) of one source (V doubles each; either may be NULL) */
int dppr_set_state(dppr_engine *e, int32_t source_index, const double *p, const double *r);
/* residual repair only (no push): the closed form checked against the sequential oracle */
int dppr_repair_only(dppr_engine *e);
/* debug: with tuning.iterlog (or DPPR_ITERLOG=1 in the environment at dppr_create), (frontier size, hub chunks, globaltimer lo, hi)
 * of every push iteration of the most recent refresh; out holds 4*cap uint32 */
int dppr_debug_iterlog(dppr_engine *e, uint32_t *out, int32_t cap, int32_t *n_out);
/* debug: 8 globaltimer stamps per CTA for push iteration DPPR_PROBE_ITER of the most recent refresh */
int dppr_debug_ctalog(dppr_engine *e, unsigned long long *out, int32_t cap_rows, int32_t *n_out);
/* stable LSD radix sort of (key, value) pairs on the device, host in / host out */
int dppr_test_sort_pairs(int32_t device, uint32_t *keys, uint32_t *vals, int64_t n, int32_t key_bits);
/* exclusive prefix sum on the device, host in / host out; returns the total in *total */
int dppr_test_exclusive_scan(int32_t device, uint32_t *data, int64_t n, uint64_t *total);
/* host-only arithmetic (no GPU needed): the internal id of the vertex of out-degree rank k among V, ranks dealt over
 * `blocks` blocks (csrc/window.cuh, relabel_slot) -- a bijection of [0, V) the CPU suite checks */
uint32_t dppr_test_relabel_slot(uint32_t rank, uint32_t vertex_count, uint32_t blocks);

#ifdef __cplusplus
}
#endif
#endif /* DPPR_H */
