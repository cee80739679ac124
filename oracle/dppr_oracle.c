/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
 *
 * Plain-C, single-threaded restatement of the reference's streaming reverse-push PPR path
 * (guowentian/dynamicppr, CPU column).  It exists so that tests, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg have a checker that travels to the GPU box (where
 * /root/reference does not exist).  Nothing under dynamicppr_b200/ may include, link or
 * execute this file; the product path fails loudly when the CUDA library is missing.
 *
 * PINNED: tests/test_oracle_golden.py checks this restatement BIT-EXACTLY (p, r, window CSR,
 * out-degrees, iteration counts) against dumps of the real reference CPU build
 * (oracle/_ref/ref_harness_serial, built by oracle/Makefile from the unmodified sources with
 * Cilk's serial elision), committed under tests/golden/ with the script that made them.
 *
 * One deliberate deviation, documented in DESIGN.md ("reference defect D1"): the reference's
 * incremental host window update (SlidingGraphVec.h:139-195) erases the FRONT k entries of a
 * vertex's adjacency vector on expiry.  For undirected streams a batch appends all original
 * orientations before all mirrored ones (SlidingGraphVec.h:266-272), so when window and batch
 * sizes are not multiples of each other the front k entries are not the k expiring ones and
 * the adjacency is transiently WRONG (the reference only notices under -DVALIDATE, which
 * rebuilds the graph from the file after every batch and thereby hides it).  This oracle
 * always uses the true window [k*B, k*B+W) in stream order -- exactly what the reference
 * produces with the line it keeps commented at cpu/PPRCPUMTCilk.h:126
 * (`dg->ConstructGraph()`), which is how the golden dumps are generated (--scratch-graph) --
 * and what the reference GPU path builds and validates (gpu/SlidingGraphBuilder.cuh:203-221,
 * gpu/PPRRevPushGPU.cuh:45-90).  For directed streams the two coincide.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_ALPHA 0.15 /* Meta.h:31 */

typedef struct DpprOracle {
    int32_t V;
    int directed;
    const int32_t *pairs; /* borrowed: M x (v1, v2), stream order */
    int64_t M, W, pos;    /* window = stream edges [pos - W, pos) */
    int32_t source;
    double eps;
    int variant;
    /* window graph, arrival order (== std::vector push_back order of
       SlidingGraphVec::ScratchConstructWindowGraph, SlidingGraphVec.h:99-137) */
    int64_t Ew;
    int32_t *in_ptr, *in_col, *out_ptr, *out_col, *deg;
    int32_t *fill; /* scratch V */
    /* app data (cpu/PPRCPUMTCilk.h:182-188, PPRCPUMTCilkRev.h:307-319) */
    double *p, *r;
    int32_t *predeg, *status;
    int32_t iteration_id;
    int32_t *ft, *ft2;
    double *ft_r;
    int32_t ft_count;
    int32_t *edge_ind;
    uint8_t *edge_flag;
    int64_t edge_cap;
    int32_t *vertex_offset;
    /* batch (EdgeBatch.h:6-30), capacity 4*B */
    int32_t *b1, *b2;
    uint8_t *bins;
    int64_t blen, bcap;
    /* 1 = reproduce reference defect D2 (DynPushInit's dedupe sentinel is `vertex_count`,
       PPRCPUMTCilkRev.h:133,141, but the values stored are offsets in [0, 2*batch_length), so
       the slot with offset == V looks "unclaimed" and its vertex can be seeded twice when
       2*batch_length > V).  Only the golden bit-exactness test turns this on. */
    int compat_d2;
    /* counters (for roofline accounting: SURVEY.md 8d) */
    int64_t n_iterations, n_pops, n_traversed;
} DpprOracle;

/* ---- A.1 window / batch arithmetic: SlidingGraphVec.h:47-66 --------------------------- */
void dppr_oracle_workload(int64_t M, double window_ratio, int mode, double batch_ratio, int64_t batch_count,
                          int64_t per_batch, int64_t total, int64_t *W_out, int64_t *B_out, int64_t *nb_out,
                          int64_t *total_out) {
    int32_t W = (int32_t)((double)(size_t)M * window_ratio); /* IndexType sliding_window_size = size_t * double */
    size_t B, nb, tot;
    if (mode == 0) { /* SLIDE_WINDOW_RATIO */
        B = (size_t)(batch_ratio * W);
        nb = (size_t)batch_count;
        tot = B * nb;
    } else { /* SLIDE_BATCH_SIZE */
        B = (size_t)per_batch;
        tot = (size_t)total;
        nb = (tot + B - 1) / B;
    }
    if (tot > (size_t)M - (size_t)W) tot = (size_t)M - (size_t)W;
    *W_out = W; *B_out = (int64_t)B; *nb_out = (int64_t)nb; *total_out = (int64_t)tot;
}

/* ---- window graph from scratch: SlidingGraphVec.h:99-137 ------------------------------ */
static void build_window(DpprOracle *o) {
    const int32_t V = o->V;
    const int64_t lo = o->pos - o->W, hi = o->pos;
    memset(o->in_ptr, 0, sizeof(int32_t) * (size_t)(V + 1));
    memset(o->out_ptr, 0, sizeof(int32_t) * (size_t)(V + 1));
    for (int64_t i = lo; i < hi; ++i) {
        int32_t a = o->pairs[2 * i], b = o->pairs[2 * i + 1];
        o->in_ptr[b + 1]++; o->out_ptr[a + 1]++;
        if (!o->directed) { o->in_ptr[a + 1]++; o->out_ptr[b + 1]++; }
    }
    for (int32_t u = 0; u < V; ++u) {
        o->in_ptr[u + 1] += o->in_ptr[u];
        o->out_ptr[u + 1] += o->out_ptr[u];
    }
    for (int32_t u = 0; u < V; ++u) o->deg[u] = o->out_ptr[u + 1] - o->out_ptr[u];
    memcpy(o->fill, o->in_ptr, sizeof(int32_t) * (size_t)V);
    for (int64_t i = lo; i < hi; ++i) {
        int32_t a = o->pairs[2 * i], b = o->pairs[2 * i + 1];
        o->in_col[o->fill[b]++] = a;
        if (!o->directed) o->in_col[o->fill[a]++] = b;
    }
    memcpy(o->fill, o->out_ptr, sizeof(int32_t) * (size_t)V);
    for (int64_t i = lo; i < hi; ++i) {
        int32_t a = o->pairs[2 * i], b = o->pairs[2 * i + 1];
        o->out_col[o->fill[a]++] = b;
        if (!o->directed) o->out_col[o->fill[b]++] = a;
    }
}

DpprOracle *dppr_oracle_create(int32_t V, int directed, const int32_t *pairs, int64_t M, int64_t W, int64_t max_batch,
                               int32_t source, double eps, int variant) {
    if (V <= 0 || W < 0 || W > M || source < 0 || source >= V || variant < 0 || variant > 3) return NULL;
    DpprOracle *o = (DpprOracle *)calloc(1, sizeof(DpprOracle));
    o->V = V; o->directed = directed ? 1 : 0; o->pairs = pairs; o->M = M; o->W = W; o->pos = W;
    o->source = source; o->eps = eps; o->variant = variant;
    o->Ew = directed ? W : 2 * W;
    size_t v1 = (size_t)V + 1, ew = (size_t)o->Ew + 1;
    o->in_ptr = (int32_t *)malloc(sizeof(int32_t) * v1);
    o->out_ptr = (int32_t *)malloc(sizeof(int32_t) * v1);
    o->in_col = (int32_t *)malloc(sizeof(int32_t) * ew);
    o->out_col = (int32_t *)malloc(sizeof(int32_t) * ew);
    o->deg = (int32_t *)malloc(sizeof(int32_t) * v1);
    o->fill = (int32_t *)malloc(sizeof(int32_t) * v1);
    o->p = (double *)malloc(sizeof(double) * v1);
    o->r = (double *)malloc(sizeof(double) * v1);
    o->predeg = (int32_t *)malloc(sizeof(int32_t) * v1);
    o->status = (int32_t *)malloc(sizeof(int32_t) * v1);
    memset(o->status, -1, sizeof(int32_t) * v1); /* PPRCPUMTCilkRev.h:13 */
    o->ft = (int32_t *)malloc(sizeof(int32_t) * v1);
    o->ft2 = (int32_t *)malloc(sizeof(int32_t) * v1);
    o->ft_r = (double *)malloc(sizeof(double) * v1);
    o->vertex_offset = (int32_t *)malloc(sizeof(int32_t) * v1);
    o->bcap = 4 * (max_batch > 0 ? max_batch : 1);
    /* edge_ind / edge_flag double as DynPushInit's scratch of 2*batch_length entries */
    o->edge_cap = (int64_t)ew;
    if (o->edge_cap < 2 * o->bcap + 1) o->edge_cap = 2 * o->bcap + 1;
    if (o->edge_cap < (int64_t)v1) o->edge_cap = (int64_t)v1;
    o->edge_ind = (int32_t *)malloc(sizeof(int32_t) * (size_t)o->edge_cap);
    o->edge_flag = (uint8_t *)malloc((size_t)o->edge_cap);
    o->b1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)o->bcap);
    o->b2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)o->bcap);
    o->bins = (uint8_t *)malloc((size_t)o->bcap);
    build_window(o);
    return o;
}

void dppr_oracle_destroy(DpprOracle *o) {
    if (!o) return;
    free(o->in_ptr); free(o->out_ptr); free(o->in_col); free(o->out_col); free(o->deg); free(o->fill);
    free(o->p); free(o->r); free(o->predeg); free(o->status); free(o->ft); free(o->ft2); free(o->ft_r);
    free(o->vertex_offset); free(o->edge_ind); free(o->edge_flag); free(o->b1); free(o->b2); free(o->bins);
    free(o);
}

/* PPRCommon.cuh:6-11 / PPRCPUMTCilkRev.h:75-80: strict inequalities */
static inline int is_legal(double r, int phase, double eps) {
    return (phase == 0 && r > eps) || (phase == 1 && r < -eps);
}

/* sequence::pack (cpu/CilkUtil.h): stable compaction of flagged entries */
static int32_t pack(const int32_t *ind, int32_t *out, const uint8_t *flag, int64_t n) {
    int32_t k = 0;
    for (int64_t i = 0; i < n; ++i)
        if (flag[i]) out[k++] = ind[i];
    return k;
}

/* One push loop to exhaustion.  Variant 0: PPRCPUMTCilkRev.h:184-289; variants 1,2,3:
 * PPRCPUMTCilkRevVariants.h (FF :231-318, Eager :112-229, Vanilla :6-110).  Serial elision of
 * cilk_for, so every "atomic" is a plain read-modify-write in loop order. */
static void main_loop(DpprOracle *o, int phase) {
    const double eps = o->eps;
    const int eager = (o->variant == 0 || o->variant == 2);     /* live residual + repair pass */
    const int status_dedupe = (o->variant == 2 || o->variant == 3);
    while (1) {
        const int32_t n = o->ft_count;
        if (n == 0) break;
        o->n_iterations++;
        o->n_pops += n;
        int32_t total = 0;
        for (int32_t i = 0; i < n; ++i) {
            int32_t u = o->ft[i];
            if (!eager) { /* snapshot + zero first (FF/Vanilla) */
                o->ft_r[i] = o->r[u];
                o->p[u] += ORACLE_ALPHA * o->r[u];
                o->r[u] = 0.0;
            }
            if (o->variant == 2) o->status[u] = o->iteration_id;
            o->vertex_offset[i] = total; /* sequence::plusScan, exclusive */
            total += o->in_ptr[u + 1] - o->in_ptr[u];
        }
        for (int32_t i = 0; i < n; ++i) {
            int32_t u = o->ft[i];
            double ru;
            if (eager) {
                ru = o->r[u];
                o->ft_r[i] = ru;
                o->p[u] += ORACLE_ALPHA * ru;
            } else {
                ru = o->ft_r[i];
            }
            const int32_t lo = o->in_ptr[u], hi = o->in_ptr[u + 1];
            o->n_traversed += hi - lo;
            for (int32_t j = lo; j < hi; ++j) {
                int64_t off = (int64_t)o->vertex_offset[i] + (j - lo);
                int32_t v = o->in_col[j];
                double add = (1.0 - ORACLE_ALPHA) * ru / (o->deg[v] + 1);
                double prer = o->r[v];
                o->r[v] = prer + add;
                double curr = prer + add;
                int is_frontier = 0;
                if (status_dedupe) {
                    if (is_legal(curr, phase, eps)) {
                        if (o->status[v] < o->iteration_id) { o->status[v] = o->iteration_id; is_frontier = 1; }
                    }
                } else {
                    if (!is_legal(prer, phase, eps) && is_legal(curr, phase, eps)) is_frontier = 1;
                }
                if (is_frontier) o->edge_ind[off] = v;
                o->edge_flag[off] = (uint8_t)is_frontier;
            }
        }
        int32_t c1 = pack(o->edge_ind, o->ft2, o->edge_flag, total);
        int32_t c2 = 0;
        if (eager) { /* repair pass: r[u] -= ru; still legal -> stays in the frontier */
            for (int32_t i = 0; i < n; ++i) {
                int32_t u = o->ft[i];
                o->r[u] -= o->ft_r[i];
                o->edge_flag[i] = (uint8_t)is_legal(o->r[u], phase, eps);
                o->edge_ind[i] = u;
            }
            c2 = pack(o->edge_ind, o->ft2 + c1, o->edge_flag, n);
        }
        int32_t *t = o->ft; o->ft = o->ft2; o->ft2 = t;
        o->ft_count = c1 + c2;
        ++o->iteration_id;
    }
}

/* PPRCPUMTCilkRev.h:38-41,174-181: Init + MainLoop(0) on the initial window */
void dppr_oracle_initial_solve(DpprOracle *o) {
    for (int32_t u = 0; u < o->V; ++u) {
        o->p[u] = 0.0;
        o->r[u] = (o->source == u) ? 1.0 : 0.0;
    }
    o->ft[0] = o->source;
    o->ft_count = 1;
    main_loop(o, 0);
}

/* DynPushInit, PPRCPUMTCilkRev.h:126-156: seed from batch endpoints, first occurrence wins */
static void dyn_push_init(DpprOracle *o, int phase) {
    const int64_t n = o->blen;
    const int32_t *set[2] = {o->b1, o->b2};
    int32_t *vmap = o->predeg; /* the reference reuses predeg as scratch (:62) */
    const int32_t unclaimed = o->compat_d2 ? o->V : INT32_MAX;
    for (int s = 0; s < 2; ++s)
        for (int64_t j = 0; j < n; ++j) vmap[set[s][j]] = unclaimed;
    for (int s = 0; s < 2; ++s)
        for (int64_t j = 0; j < n; ++j) {
            int32_t u = set[s][j];
            int64_t off = s * n + j;
            if (is_legal(o->r[u], phase, o->eps) && vmap[u] == unclaimed) {
                vmap[u] = (int32_t)off;
                o->edge_flag[off] = 1;
                o->edge_ind[off] = u;
            } else {
                o->edge_flag[off] = 0;
            }
        }
    o->ft_count = pack(o->edge_ind, o->ft, o->edge_flag, 2 * n);
}

/* One batch: SlidingGraphVec::StreamUpdates (SlidingGraphVec.h:219-275) + true-window graph +
 * IncExecuteImpl (PPRCPUMTCilkRev.h:43-73).  Returns 1 when fewer than B edges remain (:221). */
int dppr_oracle_slide(DpprOracle *o, int64_t B) {
    if (o->pos + B > o->M) return 1;
    if (4 * B > o->bcap || B > o->W) return -1;
    const int64_t left = o->pos - o->W;
    for (int64_t i = 0; i < B; ++i) { /* expiring edges first ... */
        o->b1[i] = o->pairs[2 * (left + i)]; o->b2[i] = o->pairs[2 * (left + i) + 1]; o->bins[i] = 0;
    }
    for (int64_t i = 0; i < B; ++i) { /* ... then the arriving ones */
        o->b1[B + i] = o->pairs[2 * (o->pos + i)]; o->b2[B + i] = o->pairs[2 * (o->pos + i) + 1]; o->bins[B + i] = 1;
    }
    o->blen = 2 * B;
    if (!o->directed) { /* mirrored copy appended (:266-272) */
        int64_t len = o->blen;
        memcpy(o->b1 + len, o->b2, sizeof(int32_t) * (size_t)len);
        memcpy(o->b2 + len, o->b1, sizeof(int32_t) * (size_t)len);
        memcpy(o->bins + len, o->bins, (size_t)len);
        o->blen = 2 * len;
    }
    o->pos += B;
    build_window(o);

    /* CopyOutDegree / RevertOutDegree: cpu/PPRCPUMTCilk.h:157-174 */
    for (int64_t i = 0; i < o->blen; ++i) {
        o->predeg[o->b1[i]] = o->deg[o->b1[i]];
        o->predeg[o->b2[i]] = o->deg[o->b2[i]];
    }
    for (int64_t i = 0; i < o->blen; ++i) {
        if (o->bins[i]) o->predeg[o->b1[i]]--; else o->predeg[o->b1[i]]++;
    }
    /* StreamUpdateAppData: PPRCPUMTCilkRev.h:108-124 (== gpu/StreamUpdate.cuh:35-76) */
    for (int64_t i = 0; i < o->blen; ++i) {
        int32_t u = o->b1[i], v = o->b2[i];
        double add = (1.0 - ORACLE_ALPHA) * o->p[v] - o->p[u] - ORACLE_ALPHA * o->r[u] +
                     ORACLE_ALPHA * (o->source == u ? 1.0 : 0.0);
        if (o->bins[i]) {
            o->predeg[u]++;
            o->r[u] += add / (o->predeg[u] + 1) / ORACLE_ALPHA;
        } else {
            o->predeg[u]--;
            o->r[u] -= add / (o->predeg[u] + 1) / ORACLE_ALPHA;
        }
    }
    ++o->iteration_id;
    dyn_push_init(o, 0);
    main_loop(o, 0);
    ++o->iteration_id;
    dyn_push_init(o, 1);
    main_loop(o, 1);
    return 0;
}

/* ---- accessors ------------------------------------------------------------------------ */
void dppr_oracle_set_compat_d2(DpprOracle *o, int on) { o->compat_d2 = on ? 1 : 0; }
const double *dppr_oracle_p(const DpprOracle *o) { return o->p; }
const double *dppr_oracle_r(const DpprOracle *o) { return o->r; }
const int32_t *dppr_oracle_outdeg(const DpprOracle *o) { return o->deg; }
int32_t dppr_oracle_iteration_id(const DpprOracle *o) { return o->iteration_id; }
int64_t dppr_oracle_pos(const DpprOracle *o) { return o->pos; }
void dppr_oracle_counters(const DpprOracle *o, int64_t *iters, int64_t *pops, int64_t *traversed) {
    *iters = o->n_iterations; *pops = o->n_pops; *traversed = o->n_traversed;
}

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* Canonical window graph: in_row_ptr[V+1], in_col_ind rows ascending with
 * duplicates kept, out-degree per vertex -- the object gpu/PPRRevPushGPU.cuh:45-90 compares. */
void dppr_oracle_canonical_csr(const DpprOracle *o, int32_t *in_row_ptr, int32_t *in_col_ind, int32_t *out_deg) {
    memcpy(in_row_ptr, o->in_ptr, sizeof(int32_t) * (size_t)(o->V + 1));
    memcpy(in_col_ind, o->in_col, sizeof(int32_t) * (size_t)o->Ew);
    for (int32_t u = 0; u < o->V; ++u)
        qsort(in_col_ind + in_row_ptr[u], (size_t)(in_row_ptr[u + 1] - in_row_ptr[u]), sizeof(int32_t), cmp_i32);
    memcpy(out_deg, o->deg, sizeof(int32_t) * (size_t)o->V);
}

/* Power-iteration validator, PPRCPUPowVec::CalPPRRev (cpu/PPRCPUPowVec.h:55-83): Jacobi sweeps
 * pi[u] = alpha*[u==s] + (1-alpha) * sum_{v in out(u)} pi[v] / (outdeg(u)+1) until no entry
 * moved by more than 1e-14; returns the vector BEFORE the last sweep, as the reference does. */
int dppr_oracle_power_iteration(const DpprOracle *o, double *out) {
    const int32_t V = o->V;
    double *pr[2];
    pr[0] = (double *)malloc(sizeof(double) * (size_t)V);
    pr[1] = (double *)malloc(sizeof(double) * (size_t)V);
    for (int32_t u = 0; u < V; ++u) pr[0][u] = (o->source == u) ? 1 : 0;
    int id = 0, iters = 0;
    while (1) {
        int stop = 1, oid = 1 - id;
        for (int32_t u = 0; u < V; ++u) {
            double acc = 0.0;
            int32_t lo = o->out_ptr[u], hi = o->out_ptr[u + 1];
            for (int32_t j = lo; j < hi; ++j) acc += pr[id][o->out_col[j]] / (double)((size_t)(hi - lo) + 1);
            acc = (1.0 - ORACLE_ALPHA) * acc;
            if (u == o->source) acc += ORACLE_ALPHA * 1.0;
            pr[oid][u] = acc;
            if (fabs(pr[oid][u] - pr[id][u]) > 1e-14) stop = 0;
        }
        if (stop) break;
        id = (id + 1) % 2;
        ++iters;
    }
    memcpy(out, pr[id], sizeof(double) * (size_t)V);
    free(pr[0]); free(pr[1]);
    return iters;
}

/* Residual repair alone (sequential form) on caller-provided state: used by the
 * adversarial-batch test (T7) to check the product's closed form.  deg[] must hold the
 * PRE-batch out-degrees of every u and is advanced to the post-batch values. */
void dppr_oracle_repair_sequential(int64_t n, const int32_t *u_arr, const int32_t *v_arr, const uint8_t *is_insert,
                                   int32_t source, const double *p, double *r, int32_t *deg) {
    for (int64_t i = 0; i < n; ++i) {
        int32_t u = u_arr[i], v = v_arr[i];
        double add = (1.0 - ORACLE_ALPHA) * p[v] - p[u] - ORACLE_ALPHA * r[u] + ORACLE_ALPHA * (source == u ? 1.0 : 0.0);
        if (is_insert[i]) { deg[u]++; r[u] += add / (deg[u] + 1) / ORACLE_ALPHA; }
        else              { deg[u]--; r[u] -= add / (deg[u] + 1) / ORACLE_ALPHA; }
    }
}
