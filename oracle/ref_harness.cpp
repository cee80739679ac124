// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT PATH.
//
// Oracle harness over the UNMODIFIED reference CPU implementation.  This file
// contains no PPR arithmetic of its own: it #includes the reference headers
// from /root/reference (never copied into this repo), replays the loop of
// cpu/PPRCPUMTCilk.h:101-145 (StreamUpdates -> IncConstructWindowGraph ->
// IncExecuteImpl) and dumps, after the initial solve and after every batch, the
// public members the reference itself never prints:
//   ppr->pagerank, ppr->residual           (cpu/PPRCPUMTCilk.h:182-183)
//   graph->deg, graph->in_col_ind          (GraphVec.h:128-130)
// plus (optionally) the reference's own power-iteration validator
// PPRCPUPowVec::CalPPRRev (cpu/PPRCPUPowVec.h:55-83) on the scratch-built window.
//
// Built only into oracle/_ref/ (git-ignored) by oracle/Makefile.  Used by
// tests/golden/make_golden.py (fixture generation, in the build container) and
// by tests that pin oracle/dppr_oracle.c.  Nothing under dynamicppr_b200/ may
// link or execute it.
//
// Extra flags on top of the reference CLI (Arguments.h:66-86):
//   --dump <file>     binary dump (format below)
//   --pow             also dump CalPPRRev of the scratch window per snapshot
//   --scratch-graph   call graph->ConstructGraph() after IncConstructWindowGraph
//                     (the line the reference keeps commented at
//                     cpu/PPRCPUMTCilk.h:126) so PPR always runs on the true window
//   --quiet           silence the reference's std::cout chatter
//   --times <file>    one line per (batch, source): "<batch> <ppr_us> <iteration_id> <source_index>" (cheap; used by
//                     bench.py's reference arm to drop warm-up batches -- the reference itself only prints a mean)
//   --sources a,b,c   several sources over ONE stream: one reference PPR object per source (the constructor reads
//                     gSourceVertexId, cpu/PPRCPUMTCilk.h:34), all on the same SlidingGraphVec; per batch the graph
//                     is advanced once and IncExecuteImpl runs for each source in turn, each with every thread.
//                     The reference itself runs one process per source (scripts/cpu.sh); this only shares the
//                     untimed graph maintenance.  --dump / --pow use the first source.
//
// Dump format (native little-endian):
//   char[8] "DPPRDMP1"; int32 V; int32 directed; int64 W; int64 B; int32 has_pow; int32 nsnap
//   nsnap x { int32 batch_index (0 = initial solve); int32 iteration_id;
//             int64 E; double ppr_us; int32 inc_rows_differ;
//             double p[V]; double r[V]; int32 outdeg[V];
//             int32 in_row_ptr[V+1]; int32 in_col[E]   (scratch window, rows ascending:
//                                                       the object gpu/PPRRevPushGPU.cuh:45-90 checks)
//             if has_pow: double pow[V] }
#include "Meta.h"
#include "GraphVec.h"
#include "Profiler.h"
#include "PPRCPUPowVec.h"
#include "SlidingGraphVec.h"
#include "PPRCPUMTCilkRev.h"
#include "PPRCPUMTCilkRevVariants.h"
#include "Arguments.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

namespace {

struct Snapshot {
    int32_t batch_index, iteration_id;
    int64_t E;
    double ppr_us;
    int32_t inc_rows_differ;
};

// window [file_pos/8 - W, file_pos/8) rebuilt from the file, independent of the
// reference's adjacency vectors; rows sorted ascending.
void ScratchWindowCSR(SlidingGraphVec *dg, std::vector<int32_t> &row_ptr, std::vector<int32_t> &col,
                      std::vector<int32_t> &outdeg) {
    const int V = dg->vertex_count;
    const size_t W = dg->sliding_window_size;
    std::vector<std::vector<int32_t> > in(V);
    outdeg.assign(V, 0);
    FILE *f = fopen(dg->filename.c_str(), "rb");
    assert(f != NULL);
    size_t left = dg->file_pos - W * sizeof(IndexType) * 2;
    fseek(f, left, SEEK_SET);
    std::vector<int32_t> buf(2 * W);
    size_t got = fread(buf.data(), sizeof(int32_t), 2 * W, f);
    assert(got == 2 * W);
    fclose(f);
    for (size_t i = 0; i < W; ++i) {
        int32_t a = buf[2 * i], b = buf[2 * i + 1];
        in[b].push_back(a);
        outdeg[a]++;
        if (!dg->directed) {
            in[a].push_back(b);
            outdeg[b]++;
        }
    }
    row_ptr.assign(V + 1, 0);
    col.clear();
    for (int u = 0; u < V; ++u) {
        std::sort(in[u].begin(), in[u].end());
        row_ptr[u] = (int32_t)col.size();
        col.insert(col.end(), in[u].begin(), in[u].end());
    }
    row_ptr[V] = (int32_t)col.size();
}

int CountIncRowsDiffer(SlidingGraphVec *dg, const std::vector<int32_t> &row_ptr, const std::vector<int32_t> &col,
                       const std::vector<int32_t> &outdeg) {
    int bad = 0;
    for (int u = 0; u < dg->vertex_count; ++u) {
        std::vector<int32_t> row(dg->in_col_ind[u].begin(), dg->in_col_ind[u].end());
        std::sort(row.begin(), row.end());
        bool same = (int)row.size() == row_ptr[u + 1] - row_ptr[u] && dg->deg[u] == outdeg[u];
        if (same) same = std::equal(row.begin(), row.end(), col.begin() + row_ptr[u]);
        if (!same) ++bad;
    }
    return bad;
}

}  // namespace

int main(int argc, char *argv[]) {
    std::string dump_path, times_path, sources_arg;
    bool want_pow = false, scratch_graph = false, quiet = false;
    for (int i = 1; i < argc; ++i) {
        std::string a(argv[i]);
        if (a == "--dump" && i + 1 < argc) dump_path = argv[i + 1];
        if (a == "--times" && i + 1 < argc) times_path = argv[i + 1];
        if (a == "--sources" && i + 1 < argc) sources_arg = argv[i + 1];
        if (a == "--pow") want_pow = true;
        if (a == "--scratch-graph") scratch_graph = true;
        if (a == "--quiet") quiet = true;
    }
    ArgumentsParser(argc, argv);
    std::streambuf *saved = std::cout.rdbuf();
    if (quiet) std::cout.rdbuf(NULL);
    PrintArguments();
    assert(gIsDynamic == 1);

    SlidingGraphVec *dg = new SlidingGraphVec(gDataFileName, gIsDirected);
    Profiler::InitProfiler(1, PROFILE_PHASE_NUM, PROFILE_COUNT_TYPE_NUM);

    std::vector<int> source_ids;
    for (size_t pos = 0; pos < sources_arg.size();) {
        size_t comma = sources_arg.find(',', pos);
        if (comma == std::string::npos) comma = sources_arg.size();
        if (comma > pos) source_ids.push_back(atoi(sources_arg.substr(pos, comma - pos).c_str()));
        pos = comma + 1;
    }
    if (source_ids.empty()) source_ids.push_back(gSourceVertexId);
    std::vector<PPRCPUMTCilkRev *> pprs;
    for (size_t si = 0; si < source_ids.size(); ++si) {
        gSourceVertexId = source_ids[si];
        PPRCPUMTCilkRev *one = NULL;
        if (gVariant == OPTIMIZED) one = new PPRCPUMTCilkRev(dg);
        else if (gVariant == FAST_FRONTIER) one = new PPRCPUMTCilkRevFF(dg);
        else if (gVariant == EAGER) one = new PPRCPUMTCilkRevEager(dg);
        else if (gVariant == VANILLA) one = new PPRCPUMTCilkRevVanilla(dg);
        assert(one != NULL);
        pprs.push_back(one);
    }
    PPRCPUMTCilkRev *ppr = pprs[0];

    const int V = dg->vertex_count;
    FILE *out = NULL;
    long nsnap_pos = 0;
    if (!dump_path.empty()) {
        out = fopen(dump_path.c_str(), "wb");
        assert(out != NULL);
        int32_t v = V, d = dg->directed ? 1 : 0, hp = want_pow ? 1 : 0, ns = 0;
        int64_t W = dg->sliding_window_size, B = (int64_t)gStreamUpdateCountPerBatch;
        fwrite("DPPRDMP1", 1, 8, out);
        fwrite(&v, 4, 1, out); fwrite(&d, 4, 1, out);
        fwrite(&W, 8, 1, out); fwrite(&B, 8, 1, out);
        fwrite(&hp, 4, 1, out);
        nsnap_pos = ftell(out);
        fwrite(&ns, 4, 1, out);
    }

    int32_t nsnap = 0;
    std::vector<int32_t> row_ptr, col, outdeg;
    auto snapshot = [&](int batch_index, double ppr_us, int inc_rows_differ) {
        if (!out) return;
        Snapshot s;
        s.batch_index = batch_index; s.iteration_id = ppr->iteration_id;
        s.E = row_ptr[V]; s.ppr_us = ppr_us; s.inc_rows_differ = inc_rows_differ;
        fwrite(&s.batch_index, 4, 1, out); fwrite(&s.iteration_id, 4, 1, out);
        fwrite(&s.E, 8, 1, out); fwrite(&s.ppr_us, 8, 1, out); fwrite(&s.inc_rows_differ, 4, 1, out);
        fwrite(ppr->pagerank, 8, V, out);
        fwrite(ppr->residual, 8, V, out);
        fwrite(outdeg.data(), 4, V, out);
        fwrite(row_ptr.data(), 4, V + 1, out);
        fwrite(col.data(), 4, col.size(), out);
        if (want_pow) {
            // CalPPRRev reads graph->col_ind; make sure it is the true window.
            // Save / restore the reference's incremental adjacency so that the
            // dump does not perturb the run unless --scratch-graph asked for it.
            std::vector<std::vector<IndexType> > sc = dg->col_ind, si = dg->in_col_ind;
            std::vector<IndexType> sd = dg->deg;
            dg->ConstructGraph();
            PPRCPUPowVec pw(dg);
            pw.CalPPRRev(ppr->source_vertex_id);
            fwrite(pw.pagerank, 8, V, out);
            dg->col_ind.swap(sc); dg->in_col_ind.swap(si); dg->deg.swap(sd);
        }
        ++nsnap;
    };

    FILE *times = times_path.empty() ? NULL : fopen(times_path.c_str(), "w");
    // initial solve on the first window (cpu/PPRCPUMTCilk.h:74-91)
    TimeMeasurer t0;
    for (size_t si = pprs.size(); si-- > 0;) {  // (source 0 last: t0 below stays its time)
        t0.StartTimer();
        pprs[si]->ExecuteImpl();
        t0.EndTimer();
        if (times) fprintf(times, "0 %lld %d %d\n", (long long)t0.GetElapsedMicroSeconds(), (int)pprs[si]->iteration_id, (int)si);
        if (times) fflush(times);  // a caller with a deadline reads what has been timed so far
    }
    if (out) {
        ScratchWindowCSR(dg, row_ptr, col, outdeg);
        snapshot(0, (double)t0.GetElapsedMicroSeconds(), CountIncRowsDiffer(dg, row_ptr, col, outdeg));
    }

    // streaming loop (cpu/PPRCPUMTCilk.h:101-145)
    size_t stream_batch_count = 0;
    double ppr_time = 0;
    long long total_inc_rows_differ = 0;
    while (stream_batch_count++ < gStreamBatchCount) {
        bool over = dg->StreamUpdates(gStreamUpdateCountPerBatch);
        if (over) break;
        dg->IncConstructWindowGraph();
        int differ = 0;
        if (out) {  // the window comparison is O(E_w) per batch: only when dumping
            ScratchWindowCSR(dg, row_ptr, col, outdeg);
            differ = CountIncRowsDiffer(dg, row_ptr, col, outdeg);
        }
        total_inc_rows_differ += differ;
        if (scratch_graph) dg->ConstructGraph();
        TimeMeasurer timer;
        for (size_t si = pprs.size(); si-- > 0;) {
            timer.StartTimer();
            pprs[si]->IncExecuteImpl();
            timer.EndTimer();
            ppr_time += timer.GetElapsedMicroSeconds();
            if (times) fprintf(times, "%d %lld %d %d\n", (int)stream_batch_count, (long long)timer.GetElapsedMicroSeconds(), (int)pprs[si]->iteration_id, (int)si);
            if (times) fflush(times);  // a caller with a deadline reads what has been timed so far
        }
        snapshot((int)stream_batch_count, (double)timer.GetElapsedMicroSeconds(), differ);
    }
    if (times) fclose(times);
    if (out) {
        fseek(out, nsnap_pos, SEEK_SET);
        fwrite(&nsnap, 4, 1, out);
        fclose(out);
    }
    std::cout.rdbuf(saved);
    size_t done = stream_batch_count - 1;
    double ms = ppr_time / 1000.0;
    std::cout << "harness_batches " << done << std::endl;
    std::cout << "harness_sources " << pprs.size() << std::endl;
    std::cout << "harness_inc_rows_differ " << total_inc_rows_differ << std::endl;
    std::cout << "ppr_time " << ms << " ms" << std::endl;
    std::cout << "edge_count " << (long long)gStreamUpdateCountPerBatch * (long long)done << std::endl;
    std::cout << "ppr_latency " << (done ? ms / done : 0) << " ms" << std::endl;
    std::cout << "ppr_throughput " << (ms > 0 ? gStreamUpdateCountPerBatch * done / ms * 1000.0 : 0) << " edge/s" << std::endl;
    return 0;
}
