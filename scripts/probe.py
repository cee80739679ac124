#!/usr/bin/env python
"""quick GPU probe: per-batch timings and counters of one workload (development aid, not the bench)."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamicppr_b200 import DynamicPPR, graphgen, stream, binding

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="youtube"); ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--sources", type=int, default=1); ap.add_argument("--source-kind", default="top")
ap.add_argument("--variant", type=int, default=0); ap.add_argument("--mode", type=int, default=0, help="0 auto, 1 stepwise, 2 async, 3 levelsync")
ap.add_argument("--batches", type=int, default=100); ap.add_argument("--ratio", type=float, default=0.01)
ap.add_argument("--per-batch", type=int, default=0); ap.add_argument("--eps", type=float, default=1e-9)
ap.add_argument("--hub", type=int, default=0); ap.add_argument("--show", type=int, default=5)
a = ap.parse_args()
V, M, directed = graphgen.SHAPES[a.shape]
V = max(16, int(V * a.scale)); M = max(64, int(M * a.scale))
seed = graphgen.BASE_SEED + list(graphgen.SHAPES).index(a.shape)
t0 = time.time()
edges = graphgen.rmat_directed(V, M, seed) if directed else graphgen.powerlaw_undirected(V, M, seed)
print(f"generated {a.shape} x{a.scale}: V={V} M={M} directed={directed} in {time.time()-t0:.1f}s", flush=True)
wl = stream.workload(M, 0.1, 1, -1, 0, a.per_batch, a.per_batch * a.batches) if a.per_batch else stream.workload(M, 0.1, 0, a.ratio, a.batches)
nb = wl.runnable_batches(M)
srcs = graphgen.top_out_degree(V, edges, directed, a.sources) if a.source_kind == "top" else np.arange(1, 1 + a.sources, dtype=np.int32)
print(f"W={wl.W} B={wl.B} batches={nb} sources={list(srcs[:4])}...", flush=True)
eng = DynamicPPR(V, directed, wl.W, wl.B, srcs, epsilon=a.eps, variant=a.variant, engine_mode=a.mode, hub_degree=a.hub)
t0 = time.time(); eng.init_window_pairs(edges[:wl.W]); eng.sync(); t1 = time.time()
eng.solve_initial(); eng.sync(); t2 = time.time()
s0 = eng.stats(0)
print(f"init_window {1e3*(t1-t0):.2f} ms; initial solve {s0.ms_push:.3f} ms (wall {1e3*(t2-t1):.2f}) iters={s0.iterations} pops={s0.frontier_pops} T={s0.traversed_edges}")
t0 = time.time()
for k in range(nb):
    lo = wl.W + k * wl.B
    eng.slide_pairs(edges[lo:lo + wl.B])
eng.sync(); wall = time.time() - t0
rows = [eng.stats(k + 1) for k in range(nb)]
err = max(r.error_flags for r in rows)
f = lambda name: np.array([getattr(r, name) for r in rows], dtype=np.float64)
ppr = f("ms_repair") + f("ms_push"); e2e = ppr + f("ms_window") + f("ms_upload")
print(f"errors={err} wall={1e3*wall:.1f} ms for {nb} batches ({1e3*wall/nb:.3f} ms/batch)")
print(f"mean ms: upload {f('ms_upload').mean():.4f} window {f('ms_window').mean():.4f} repair {f('ms_repair').mean():.4f} push {f('ms_push').mean():.4f}")
print(f"ppr p50 {np.median(ppr):.4f} ms p95 {np.percentile(ppr,95):.4f}; e2e p50 {np.median(e2e):.4f}")
print(f"ppr_throughput {wl.B*nb/ppr.sum()*1e3:.4g} edges/s; e2e {wl.B*nb/e2e.sum()*1e3:.4g} edges/s; x{len(srcs)} sources")
print(f"per batch: iters {f('iterations').mean():.1f} pops {f('frontier_pops').mean():.0f} traversed {f('traversed_edges').mean():.0f} hubs {f('hub_pops').mean():.1f} reloc {f('relocations').mean():.1f} touched {f('touched_vertices').mean():.0f}")
T = f('traversed_edges').sum(); F = f('frontier_pops').sum()
print(f"push algorithmic GB/s: {(24*T+56*F)/ (f('ms_push').sum()*1e-3)/1e9:.1f}; edges/us {T/(f('ms_push').sum()*1e3):.1f}; us/iter {f('ms_push').sum()*1e3/max(f('iterations').sum(),1):.2f}")
for r in rows[:a.show]:
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.as_dict().items()})
if os.environ.get("DPPR_ITERLOG"):
    lg = eng.iterlog()
    if len(lg):
        t = lg[:, 2].astype(np.int64); dt = np.diff(t) / 1e3
        print("last batch per-iteration: (frontier, hub_chunks, us)")
        print(" ".join(f"({int(a)},{int(b)},{c:.1f})" for (a, b, _), c in zip(lg[:-1], dt)))
    cl = eng.ctalog().astype(np.int64)
    if len(cl) and cl[:, 0].max() > 0:
        t0 = cl[:, 0].min()
        rel = (cl - t0) / 1e3
        names = ["start", "hubs_done", "pop_done", "scan_done", "edges_done", "flush_done", "tiles_done", "barrier_done"]
        print(f"probe iteration {os.environ.get('DPPR_PROBE_ITER', 10)}: per-CTA stamps (us from earliest start), {len(cl)} CTAs")
        for j, nm in enumerate(names):
            col = rel[:, j][cl[:, j] > 0]
            if len(col):
                print(f"  {nm:13s} min {col.min():7.2f} p50 {np.median(col):7.2f} p90 {np.percentile(col, 90):7.2f} max {col.max():7.2f}")
