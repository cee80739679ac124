#!/usr/bin/env python
"""Run one BASELINE.json config at full size on the GPU box and print one JSON line with timings, counters and the
size-independent parity properties (window CSR vs numpy lexsort, residual bound, push invariant).  Results go to
profiles/.  (Configs 3-5 are too large for the pytest suite.)"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dynamicppr_b200 import DynamicPPR, graphgen, stream
from test_gpu_fullsize import numpy_window_csr, invariant_defect

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True, help="3 = LiveJournal-shaped mode 1, 4 = Orkut-shaped multi-source")
ap.add_argument("--sources", type=int, default=1); ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--batches", type=int, default=100); ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--check", type=int, default=1)
a = ap.parse_args()
shape = {3: "livejournal", 4: "orkut"}[a.config]
V, M, directed = graphgen.SHAPES[shape]
V, M = int(V * a.scale), int(M * a.scale)
seed = graphgen.BASE_SEED + list(graphgen.SHAPES).index(shape)
t0 = time.time()
edges = graphgen.rmat_directed(V, M, seed) if directed else graphgen.powerlaw_undirected(V, M, seed)
tgen = time.time() - t0
wl = stream.workload(M, 0.1, 1, -1, 0, 100, 100 * a.batches) if a.config == 3 else stream.workload(M, 0.1, 0, 0.01, a.batches)
nb = min(a.batches, wl.runnable_batches(M))
srcs = graphgen.top_out_degree(V, edges, directed, a.sources)
eps = 1e-9
t0 = time.time()
eng = DynamicPPR(V, directed, wl.W, wl.B, srcs, epsilon=eps, variant=a.variant, frontier_capacity=min(V * len(srcs), 1 << 30))
eng.init_window_pairs(edges[:wl.W]); eng.sync(); tinit = time.time() - t0
eng.solve_initial(); eng.sync(); s0 = eng.stats(0)
for k in range(nb):
    lo = wl.W + k * wl.B
    eng.slide_pairs(edges[lo:lo + wl.B])
eng.sync()
rows = [eng.stats(k + 1) for k in range(nb)]
f = lambda n: np.array([getattr(r, n) for r in rows], dtype=np.float64)
ppr = f("ms_repair") + f("ms_push"); e2e = ppr + f("ms_window") + f("ms_upload")
T, F = f("traversed_edges"), f("frontier_pops")
out = dict(config=a.config, shape=shape, V=V, M=M, W=wl.W, B=wl.B, batches=nb, sources=len(srcs), variant=a.variant,
           gen_s=round(tgen, 1), init_window_s=round(tinit, 2), initial_solve_ms=s0.ms_push,
           error_flags=int(max(r.error_flags for r in rows)),
           ppr_ms_mean=float(ppr.mean()), ppr_ms_p50=float(np.median(ppr)), ppr_ms_p95=float(np.percentile(ppr, 95)),
           e2e_ms_mean=float(e2e.mean()), window_ms_mean=float(f("ms_window").mean()), repair_ms_mean=float(f("ms_repair").mean()),
           edge_updates_per_s=float(wl.B * nb / ppr.sum() * 1e3), source_edge_updates_per_s=float(len(srcs) * wl.B * nb / e2e.sum() * 1e3),
           iterations=float(f("iterations").mean()), dense_sweeps=float(f("dense_sweeps").mean()), pops=float(F.mean()), traversed=float(T.mean()),
           traversed_per_update=float(T.sum() / (wl.B * nb * len(srcs))),
           push_edges_per_ns=float(T.sum() / (f("ms_push").sum() * 1e6)),
           push_alg_GBps=float((24 * T + 56 * F).sum() / (f("ms_push").sum() * 1e-3) / 1e9),
           repair_alg_GBps=float(45.0 * f("batch_entries").sum() * len(srcs) / (f("ms_repair").sum() * 1e-3) / 1e9),
           pool_used=int(rows[-1].pool_used), relocations=float(f("relocations").mean()))
if a.check:
    rp, ci, od = eng.export_window_csr()
    erp, eci, eod = numpy_window_csr(V, directed, edges[nb * wl.B: nb * wl.B + wl.W])
    out["window_bit_exact"] = bool(np.array_equal(rp, erp) and np.array_equal(ci, eci) and np.array_equal(od, eod))
    worst_r, worst_inv = 0.0, 0.0
    for i in range(min(len(srcs), 4)):
        p, r = eng.estimates(i), eng.residuals(i)
        worst_r = max(worst_r, float(np.abs(r).max()))
        worst_inv = max(worst_inv, float(invariant_defect(V, rp, ci, od, p, r, int(srcs[i]))))
    out["max_abs_residual_over_eps"] = worst_r / eps
    out["invariant_defect"] = worst_inv
print(json.dumps(out))
