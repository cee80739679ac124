// window.cuh -- device-resident sliding-window graph (north-star subsystem 1).
//
// Replaces gpu/SlidingGraphBuilder.cuh, which after EVERY batch shifts the whole arrival-order
// edge list and rebuilds the in-CSR by a comparator sort of all E_w window edges
// (:163-181, :203-221) and recounts every out-degree (:193-201).  Here the update is O(batch):
//
//   * `log`   : the W window edges in arrival order, kept as a RING (no shifting).  The B edges
//               that expire are exactly the B ring slots the arriving edges overwrite.
//   * in-adjacency: one FIFO ring per vertex inside a pooled slot array.  A window slides in
//               stream order, so per vertex the expiring in-edges are always the oldest ones:
//               expire = advance `head`, insert = append at `head+len`.  VMeta{base,head,len,cap}
//               is one 16-byte load per frontier pop; cap is a power of two (ring index = mask).
//               A ring that fills up moves to a larger slot range (amortised doubling); one that has shrunk to a
//               quarter of its capacity moves to a smaller one; the range it leaves goes onto a per-size-class free
//               stack (one batch later: the relocation copy still reads it) and is handed out again before the pool's
//               bump pointer advances.  Pool use therefore tracks the LIVE degree sum, not the sum of every vertex's
//               peak degree over the life of the stream (round-1 advisor finding).
//   * out-degree: plain int32 per vertex (the reference only ever uses row_ptr differences,
//               gpu/ExpandRev.cuh:71, gpu/StreamUpdate.cuh:13).
//   * out-adjacency (directed graphs, when dense iterations are enabled -- pull.cuh): a second set of per-vertex
//               FIFO rings keyed by the source end, in the same pool, maintained by the same plan / relocate /
//               insert steps from the batch's source-sorted entries.  Undirected graphs need none: their entry set
//               is symmetric, the in-lists are the out-lists.
//
// Determinism: the batch's directed entries are radix-sorted by vertex (stable => stream order
// within a vertex), run-length encoded, and every per-vertex quantity is then written by exactly
// one thread.  No order-dependent atomic touches the graph, which is what makes the exported
// canonical CSR bit-exact with the reference's (gpu/PPRRevPushGPU.cuh:45-90).
#pragma once
#include "common.cuh"

namespace dppr {

enum : int {
    kErrPool = 1, kErrQueue = 2, kErrHubQ = 4, kErrWatchdog = 8, kErrUnderflow = 16, kErrBadId = 32
};

struct VMeta {            // layout-compatible with uint4
    uint32_t base, head, len, cap;
};

struct RelocJob {
    uint32_t old_base, old_head, old_cap, len, new_base, pad;
};

// free ranges of the adjacency pool, one stack per power-of-two size class (class c holds ranges of 2^c slots)
struct PoolFree {
    int *count;              // [32] entries on each stack (transiently negative while several pops race for the last entry)
    const uint32_t *off;     // [33] stack c lives in stack[off[c], off[c+1])
    uint32_t *stack;
    uint2 *pend;             // ranges released by THIS batch: {base, capacity}; pushed onto the stacks by the next batch
    uint32_t *npend;
    unsigned long long *leaked;  // slots dropped because a stack was full
};

struct WindowView {
    int32_t V;
    uint4 *vmeta;
    int32_t *pool;
    int32_t *outdeg;
    unsigned long long *pool_top;
    unsigned long long pool_cap;
    int *errflags;
    PoolFree fr;
};

constexpr uint32_t kNoRange = 0xffffffffu;

// A range of `cap` (power of two) slots: the free stack of its class first, the bump pointer otherwise.  Called from the
// plan stage only, where nothing pushes: concurrent pops are safe with a plain fetch-and-subtract.
__device__ __forceinline__ uint32_t pool_alloc(const WindowView &w, uint32_t cap) {
    const int cls = 31 - __clz((int)cap);
    if (__ldcg(&w.fr.count[cls]) > 0) {
        const int i = atomicSub(&w.fr.count[cls], 1) - 1;
        if (i >= 0) return __ldcg(&w.fr.stack[w.fr.off[cls] + (uint32_t)i]);
        atomicAdd(&w.fr.count[cls], 1);  // lost the race for the last entry
    }
    const unsigned long long nb = atomicAdd(w.pool_top, (unsigned long long)cap);
    return nb + cap > w.pool_cap ? kNoRange : (uint32_t)nb;
}

// Hands the ranges the PREVIOUS batch released to the free stacks.  Runs in the first stage of a window update (entries),
// which is separated from the plan stage by a kernel boundary / grid barrier: pushes never race with pops.
__device__ __forceinline__ void pool_reclaim_one(uint32_t j, const uint2 *pend_prev, const PoolFree &fr) {
    const uint2 r = __ldcg(&pend_prev[j]);
    const int cls = 31 - __clz((int)r.y);
    const uint32_t room = fr.off[cls + 1] - fr.off[cls];
    const int i = atomicAdd(&fr.count[cls], 1);
    if (i >= 0 && (uint32_t)i < room) {
        fr.stack[fr.off[cls] + (uint32_t)i] = r.x;
    } else {
        atomicSub(&fr.count[cls], 1);
        atomicAdd(fr.leaked, (unsigned long long)r.y);
    }
}

__host__ __device__ __forceinline__ uint32_t next_pow2_u32(uint32_t x) {
    if (x <= 1) return 1;
    --x;
    x |= x >> 1; x |= x >> 2; x |= x >> 4; x |= x >> 8; x |= x >> 16;
    return x + 1;
}
// capacity policy: 12.5% headroom, at least 4 slots, power of two
__host__ __device__ __forceinline__ uint32_t ring_capacity_for(uint32_t need) {
    uint32_t c = next_pow2_u32(need + (need >> 3));
    return c < 4u ? 4u : c;
}

// ---------------------------------------------------------------------------------------------
// internal vertex order.  The engine renumbers vertices by descending out-degree of the initial window: a
// vertex receives one residual add per out-neighbour that is pushed, so the hubs' residuals and degrees -- the
// bulk of all random traffic on a power-law graph -- become a dense, cache-resident prefix of r[] / outdeg[]
// instead of being scattered over the whole array.  perm: caller id -> internal id, inv: internal -> caller.
// Ids are translated where edges enter (window init, batch entries) and where results leave (estimates,
// residuals, exported CSR); everything in between is id-agnostic.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    relabel_degrees(const int2 *__restrict__ log, int64_t W, int directed, int32_t V, uint32_t *__restrict__ deg, int *errflags,
                    int both_ends) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < W; i += (int64_t)gridDim.x * kThreads) {
        const int2 e = log[i];
        if ((uint32_t)e.x >= (uint32_t)V || (uint32_t)e.y >= (uint32_t)V) { atomicOr(errflags, kErrBadId); continue; }
        atomicAdd(&deg[e.x], 1u);
        if (!directed || both_ends) atomicAdd(&deg[e.y], 1u);  // both_ends: rank by out- plus in-degree
    }
}
__global__ void __launch_bounds__(kThreads)
    relabel_keys(const uint32_t *__restrict__ deg, uint32_t degmax, uint32_t *__restrict__ key, uint32_t *__restrict__ val, int32_t V) {
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < V; v += (int64_t)gridDim.x * kThreads) {
        key[v] = degmax - min(deg[v], degmax);  // ascending key = descending degree; the stable sort keeps id order in ties
        val[v] = (uint32_t)v;
    }
}
// rank k (0 = highest out-degree) -> internal id.  Ranks are dealt round-robin over P blocks: the hot vertices
// still fill a dense prefix of every block (what keeps them cache-resident when the arrays exceed L2), but
// CONSECUTIVE ranks -- the very hottest addresses -- land ~V/P entries apart instead of in the same 32-byte sector,
// whose atomics would otherwise serialise in one L2 slice (measured: plain rank order costs +29 % on L2-resident
// graphs, this interleaved order is neutral there and +29 % faster on the DRAM-resident Twitter-shaped graph).
__host__ __device__ __forceinline__ uint32_t relabel_slot(uint32_t k, uint32_t V, uint32_t P) {
    const uint32_t r = k % P, c = k / P, q = V / P, rem = V % P;
    return r * q + (r < rem ? r : rem) + c;
}
__global__ void __launch_bounds__(kThreads)
    relabel_assign(const uint32_t *__restrict__ by_rank, uint32_t *__restrict__ perm, uint32_t *__restrict__ inv, int32_t V,
                   uint32_t P) {
    for (int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x; k < V; k += (int64_t)gridDim.x * kThreads) {
        const uint32_t id = relabel_slot((uint32_t)k, (uint32_t)V, P);
        const uint32_t v = by_rank[k];
        inv[id] = v;
        perm[v] = id;
    }
}
__global__ void __launch_bounds__(kThreads)
    relabel_log(int2 *__restrict__ log, int64_t W, const uint32_t *__restrict__ perm, int32_t V) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < W; i += (int64_t)gridDim.x * kThreads) {
        int2 e = log[i];
        if ((uint32_t)e.x < (uint32_t)V && (uint32_t)e.y < (uint32_t)V) log[i] = make_int2((int)perm[e.x], (int)perm[e.y]);
    }
}
// out[v] = in[perm[v]] (results leave in caller ids) / out[perm[v]] = in[v] (state enters in caller ids)
template <typename T>
__global__ void __launch_bounds__(kThreads)
    gather_by_perm(const T *__restrict__ in, int64_t in_stride, const uint32_t *__restrict__ perm, T *__restrict__ out, int32_t V) {
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < V; v += (int64_t)gridDim.x * kThreads)
        out[v] = in[(int64_t)(perm ? perm[v] : (uint32_t)v) * in_stride];
}
template <typename T>
__global__ void __launch_bounds__(kThreads)
    scatter_by_perm(const T *__restrict__ in, const uint32_t *__restrict__ perm, T *__restrict__ out, int64_t out_stride, int32_t V) {
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < V; v += (int64_t)gridDim.x * kThreads)
        out[(int64_t)(perm ? perm[v] : (uint32_t)v) * out_stride] = in[v];
}

// ---------------------------------------------------------------------------------------------
// initial window: entries (dst, src) in stream order + degree histograms
// (replaces InitWindowStream + BuildInGraph, gpu/SlidingGraphBuilder.cuh:182-192,145-160)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    win_init_entries(const int2 *__restrict__ log, int64_t W, int directed, int32_t V, uint32_t *__restrict__ key,
                     uint32_t *__restrict__ val, uint32_t *__restrict__ indeg, int32_t *__restrict__ outdeg,
                     int *errflags, int transpose) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < W; i += (int64_t)gridDim.x * kThreads) {
        int2 e = log[i];
        if ((uint32_t)e.x >= (uint32_t)V || (uint32_t)e.y >= (uint32_t)V) {
            atomicOr(errflags, kErrBadId);
            e.x = 0; e.y = 0;
        }
        if (directed && transpose) {  // out-lists of a directed graph (pull.cuh): keyed by the source end
            key[i] = (uint32_t)e.x; val[i] = (uint32_t)e.y;
            atomicAdd(&indeg[e.x], 1u);
        } else if (directed) {
            key[i] = (uint32_t)e.y; val[i] = (uint32_t)e.x;
            atomicAdd(&indeg[e.y], 1u);
            atomicAdd(&outdeg[e.x], 1);
        } else {  // mirrored at load (SlidingGraphVec.h:84-90); same interleaving as the host push_back order
            key[2 * i] = (uint32_t)e.y; val[2 * i] = (uint32_t)e.x;
            key[2 * i + 1] = (uint32_t)e.x; val[2 * i + 1] = (uint32_t)e.y;
            atomicAdd(&indeg[e.y], 1u); atomicAdd(&indeg[e.x], 1u);
            atomicAdd(&outdeg[e.x], 1); atomicAdd(&outdeg[e.y], 1);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
    win_init_caps(const uint32_t *__restrict__ indeg, uint32_t *__restrict__ caps, int32_t V) {
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < V; v += (int64_t)gridDim.x * kThreads)
        caps[v] = indeg[v] ? ring_capacity_for(indeg[v]) : 0u;
}

__global__ void __launch_bounds__(kThreads)
    win_init_meta(const uint32_t *__restrict__ indeg, const uint32_t *__restrict__ caps,
                  const uint32_t *__restrict__ capbase, uint4 *__restrict__ vmeta, int32_t V, uint32_t base_off) {
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < V; v += (int64_t)gridDim.x * kThreads)
        vmeta[v] = make_uint4(capbase[v] + base_off, 0u, indeg[v], caps[v]);
}

// sorted (dst, src) entries -> ring slots; rank within the row = i - rowptr[dst]
__global__ void __launch_bounds__(kThreads)
    win_init_fill(const uint32_t *__restrict__ key, const uint32_t *__restrict__ val, int64_t n,
                  const uint32_t *__restrict__ rowptr, const uint32_t *__restrict__ capbase, int32_t *__restrict__ pool,
                  uint32_t base_off) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const uint32_t d = key[i];
        pool[base_off + capbase[d] + ((uint32_t)i - rowptr[d])] = (int32_t)val[i];
    }
}

// ---------------------------------------------------------------------------------------------
// per batch: expiring + arriving edges -> directed entries (vertex key, (other << 1) | is_insert)
// Entry order = reference EdgeBatch order (deletes then inserts, SlidingGraphVec.h:238-262) with
// the two orientations of an undirected edge interleaved in stream order.
//   directed  : group A keyed by dst (in-lists), group B keyed by src (out-degree + residual repair)
//   undirected: the entry set is symmetric, one group keyed by u serves both roles
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void batch_entries_one(int64_t i, int2 *log, int64_t W, int64_t log_start,
                                                  const int2 *arriving, int64_t B, int directed, int32_t V,
                                                  uint32_t *akey, uint32_t *aval,
                                                  uint32_t *bkey, uint32_t *bval, int *errflags, const uint32_t *perm) {
    int64_t slot = log_start + i;
    if (slot >= W) slot -= W;
    const int2 old = log[slot];  // already in internal ids
    int2 nw = arriving[i];
    if ((uint32_t)nw.x >= (uint32_t)V || (uint32_t)nw.y >= (uint32_t)V) {
        atomicOr(errflags, kErrBadId);
        nw.x = 0; nw.y = 0;
    }
    if (perm) nw = make_int2((int)perm[nw.x], (int)perm[nw.y]);
    log[slot] = nw;
    if (directed) {
        akey[i] = (uint32_t)old.y;     aval[i] = ((uint32_t)old.x << 1);
        akey[B + i] = (uint32_t)nw.y;  aval[B + i] = ((uint32_t)nw.x << 1) | 1u;
        bkey[i] = (uint32_t)old.x;     bval[i] = ((uint32_t)old.y << 1);
        bkey[B + i] = (uint32_t)nw.x;  bval[B + i] = ((uint32_t)nw.y << 1) | 1u;
    } else {
        akey[2 * i] = (uint32_t)old.y;              aval[2 * i] = ((uint32_t)old.x << 1);
        akey[2 * i + 1] = (uint32_t)old.x;          aval[2 * i + 1] = ((uint32_t)old.y << 1);
        akey[2 * B + 2 * i] = (uint32_t)nw.y;       aval[2 * B + 2 * i] = ((uint32_t)nw.x << 1) | 1u;
        akey[2 * B + 2 * i + 1] = (uint32_t)nw.x;   aval[2 * B + 2 * i + 1] = ((uint32_t)nw.y << 1) | 1u;
    }
}

__global__ void __launch_bounds__(kThreads)
    win_batch_entries(int2 *__restrict__ log, int64_t W, int64_t log_start, const int2 *__restrict__ arriving, int64_t B,
                      int directed, int32_t V, uint32_t *__restrict__ akey, uint32_t *__restrict__ aval,
                      uint32_t *__restrict__ bkey, uint32_t *__restrict__ bval, int *errflags, const uint32_t *perm,
                      const uint2 *__restrict__ pend_prev, const uint32_t *__restrict__ npend_prev, PoolFree fr) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < B; i += (int64_t)gridDim.x * kThreads)
        batch_entries_one(i, log, W, log_start, arriving, B, directed, V, akey, aval, bkey, bval, errflags, perm);
    const uint32_t np = *npend_prev;
    for (uint32_t j = blockIdx.x * kThreads + threadIdx.x; j < np; j += gridDim.x * kThreads) pool_reclaim_one(j, pend_prev, fr);
}

// ---------------------------------------------------------------------------------------------
// run-length encoding of the sorted entries.  Within a vertex's run the stable sort leaves all
// deletes before all inserts, so a run is described by (start, first_insert, end).
// ---------------------------------------------------------------------------------------------
struct Segments {
    uint32_t *vertex;      // [cap]   run -> vertex id
    uint32_t *start;       // [cap+1] run -> first entry
    uint32_t *first_ins;   // [cap]   run -> first insert entry (== end if none)
    uint32_t *segof;       // [cap]   entry -> run
    uint32_t *count;       // device scalar: number of runs
};

__global__ void __launch_bounds__(kThreads)
    rle_heads(const uint32_t *__restrict__ key, int64_t n, uint32_t *__restrict__ flag) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}

// entry i of the sorted batch belongs to run s (head = first entry of its run).
// (no __restrict__ on the *_one helpers: the fused single-CTA kernel reads what it wrote earlier in the same
// launch, so their loads must not be turned into non-coherent LDG.NC)
__device__ __forceinline__ void rle_fill_one(int64_t i, uint32_t s, bool head, const uint32_t *key,
                                             const uint32_t *val, int64_t n, const Segments &sg) {
    const uint32_t k = key[i];
    const bool last = (i == n - 1) || (key[i + 1] != k);
    const bool ins = val[i] & 1u;
    if (head) {
        sg.vertex[s] = k;
        sg.start[s] = (uint32_t)i;
    }
    if (ins && (head || !(val[i - 1] & 1u))) sg.first_ins[s] = (uint32_t)i;
    if (last && !ins) sg.first_ins[s] = (uint32_t)i + 1u;
    if (i == n - 1) {
        sg.start[s + 1] = (uint32_t)n;
        *sg.count = s + 1u;
    }
    sg.segof[i] = s;
}

// `sg.segof` holds the exclusive scan of the head flags on entry and the run index on exit.
__global__ void __launch_bounds__(kThreads)
    rle_fill(const uint32_t *__restrict__ key, const uint32_t *__restrict__ val, int64_t n, Segments sg) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const bool head = (i == 0) || (key[i - 1] != key[i]);
        rle_fill_one(i, sg.segof[i] + (head ? 1u : 0u) - 1u, head, key, val, n, sg);  // each thread rewrites only its own slot
    }
}

// one thread per touched vertex: expire (advance head), reserve room for the inserts, grow or shrink the ring
__device__ __forceinline__ void plan_one(uint32_t s, const Segments &sg, const WindowView &w, uint32_t *ins_pos,
                                         RelocJob *jobs, uint32_t *njobs) {
    const uint32_t v = sg.vertex[s];
    const uint32_t st = sg.start[s], fi = sg.first_ins[s], en = sg.start[s + 1];
    uint32_t ndel = fi - st;
    const uint32_t nins = en - fi;
    uint4 q = w.vmeta[v];
    VMeta m{q.x, q.y, q.z, q.w};
    if (ndel > m.len) {  // the caller slid edges the window never held
        atomicOr(w.errflags, kErrUnderflow);
        ndel = m.len;
    }
    if (ndel) {
        m.head = (m.head + ndel) & (m.cap - 1u);
        m.len -= ndel;
    }
    uint32_t pos = m.len;
    const uint32_t need = m.len + nins;
    const bool grow = need > m.cap;
    const bool shrink = !grow && m.cap > 4u && need <= (m.cap >> 2);  // hysteresis: it grows again only beyond 2x
    if (grow || shrink) {
        const uint32_t ncap = need ? ring_capacity_for(need) : 0u;  // an emptied ring gives its range back entirely
        const uint32_t nb = ncap ? pool_alloc(w, ncap) : 0u;
        if (nb == kNoRange) {
            if (grow) {
                atomicOr(w.errflags, kErrPool);
                pos = 0xffffffffu;  // inserts of this run are dropped; the engine is flagged and refuses further batches
            }                       // (a shrink that finds no range simply keeps the ring)
        } else {
            if (m.cap) w.fr.pend[atomicAdd(w.fr.npend, 1u)] = make_uint2(m.base, m.cap);
            if (m.len) {
                const uint32_t j = atomicAdd(njobs, 1u);
                jobs[j] = RelocJob{m.base, m.head, m.cap, m.len, nb, 0u};
            }
            m.base = nb;
            m.head = 0u;
            m.cap = ncap;
        }
    }
    if (nins && pos != 0xffffffffu) m.len += nins;
    ins_pos[s] = pos;
    w.vmeta[v] = make_uint4(m.base, m.head, m.len, m.cap);
}

__global__ void __launch_bounds__(kThreads)
    win_plan(Segments sg, WindowView w, uint32_t *__restrict__ ins_pos, RelocJob *__restrict__ jobs, uint32_t *njobs) {
    const uint32_t nseg = *sg.count;
    for (uint32_t s = blockIdx.x * kThreads + threadIdx.x; s < nseg; s += gridDim.x * kThreads)
        plan_one(s, sg, w, ins_pos, jobs, njobs);
}

// one CTA per relocated ring (grid-stride over jobs)
__global__ void __launch_bounds__(kThreads)
    win_relocate(const RelocJob *__restrict__ jobs, const uint32_t *__restrict__ njobs, int32_t *__restrict__ pool) {
    const uint32_t n = *njobs;
    for (uint32_t j = blockIdx.x; j < n; j += gridDim.x) {
        const RelocJob jb = jobs[j];
        for (uint32_t k = threadIdx.x; k < jb.len; k += kThreads)
            pool[jb.new_base + k] = pool[jb.old_base + ((jb.old_head + k) & (jb.old_cap - 1u))];
    }
}

// one thread per entry: inserts take slot head + (len before inserts) + rank within the run
__device__ __forceinline__ void insert_one(int64_t i, const uint32_t *key, const uint32_t *val,
                                           const Segments &sg, const uint32_t *ins_pos, const WindowView &w) {
    const uint32_t x = val[i];
    if (!(x & 1u)) return;
    const uint32_t s = sg.segof[i];
    const uint32_t pos = ins_pos[s];
    if (pos == 0xffffffffu) return;
    const uint4 m = w.vmeta[key[i]];
    const uint32_t rank = (uint32_t)i - sg.first_ins[s];
    w.pool[m.x + ((m.y + pos + rank) & (m.w - 1u))] = (int32_t)(x >> 1);
}

__global__ void __launch_bounds__(kThreads)
    win_insert(const uint32_t *__restrict__ key, const uint32_t *__restrict__ val, int64_t n, Segments sg,
               const uint32_t *__restrict__ ins_pos, WindowView w) {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        insert_one(i, key, val, sg, ins_pos, w);
}

// out-degree side (group keyed by src): remember the pre-batch degree for the repair, store the new one
__device__ __forceinline__ void out_degree_one(uint32_t s, const Segments &sg, int32_t *outdeg,
                                               int32_t *seg_d0) {
    const uint32_t u = sg.vertex[s];
    const uint32_t st = sg.start[s], fi = sg.first_ins[s], en = sg.start[s + 1];
    const int32_t d0 = outdeg[u];
    seg_d0[s] = d0;
    outdeg[u] = d0 + (int32_t)(en - fi) - (int32_t)(fi - st);
}

__global__ void __launch_bounds__(kThreads)
    win_out_degrees(Segments sg, int32_t *__restrict__ outdeg, int32_t *__restrict__ seg_d0) {
    const uint32_t nseg = *sg.count;
    for (uint32_t s = blockIdx.x * kThreads + threadIdx.x; s < nseg; s += gridDim.x * kThreads)
        out_degree_one(s, sg, outdeg, seg_d0);
}

// ---------------------------------------------------------------------------------------------
// export (test / validation path): ring contents -> (dst, src) entries
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    win_export_len(const uint4 *__restrict__ vmeta, uint32_t *__restrict__ len, int32_t V, const uint32_t *__restrict__ perm) {
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < V; v += (int64_t)gridDim.x * kThreads)
        len[v] = vmeta[perm ? perm[v] : (uint32_t)v].z;
}

// one warp per vertex (caller ids on both sides of every exported entry)
__global__ void __launch_bounds__(kThreads)
    win_export_entries(const uint4 *__restrict__ vmeta, const int32_t *__restrict__ pool,
                       const uint32_t *__restrict__ rowptr, uint32_t *__restrict__ key, uint32_t *__restrict__ val,
                       int32_t V, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ inv) {
    const int64_t warps = (int64_t)gridDim.x * kWarps;
    for (int64_t v = (int64_t)blockIdx.x * kWarps + warp_id(); v < V; v += warps) {
        const uint4 m = vmeta[perm ? perm[v] : (uint32_t)v];
        const uint32_t o = rowptr[v];
        for (uint32_t k = lane_id(); k < m.z; k += 32) {
            const uint32_t nb = (uint32_t)pool[m.x + ((m.y + k) & (m.w - 1u))];
            key[o + k] = (uint32_t)v;
            val[o + k] = inv ? inv[nb] : nb;
        }
    }
}

}  // namespace dppr
