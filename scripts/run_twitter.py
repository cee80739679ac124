#!/usr/bin/env python
"""BASELINE config 5 shape on one B200: Twitter-2010-shaped directed R-MAT stream (41,652,230 V, 1,468,365,182 E),
window 0.1 (146,836,518 edges), -r 0.01 (1,468,365 edges per batch), eps 1e-9.  The 11.7 GB stream is generated ON the
device (dppr_generate_rmat_device) and never touches the host; the engine is fed device pointers.  Prints one JSON
line per source kind with timings, counters and the size-independent parity properties."""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from dynamicppr_b200 import DynamicPPR, binding, stream
from test_gpu_fullsize import invariant_defect

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0); ap.add_argument("--batches", type=int, default=20)
ap.add_argument("--kinds", default="low,top"); ap.add_argument("--check", type=int, default=1)
ap.add_argument("--sources", type=int, default=1); ap.add_argument("--top-batches", type=int, default=3)
ap.add_argument("--V", type=int, default=41_652_230); ap.add_argument("--M", type=int, default=1_468_365_182)
ap.add_argument("--undirected", type=int, default=0, help="1: feed the same device-generated pairs as an UNDIRECTED stream (perf exploration of other shapes)")
a = ap.parse_args()
V, M = int(a.V * a.scale), int(a.M * a.scale)
wl = stream.workload(M, 0.1, 0, 0.01, a.batches)
nb = min(a.batches, wl.runnable_batches(M))
need = wl.W + nb * wl.B
t0 = time.time()
dev = torch.empty((need, 2), dtype=torch.int32, device="cuda")
binding.generate_rmat_device(V, need, 20261021, dev.data_ptr())
torch.cuda.synchronize(); tgen = time.time() - t0
outdeg_all = torch.bincount(dev[: wl.W, 0].long(), minlength=V)  # (initial window only: the choice must not depend on --batches)
# source buckets of the reference's workload tool (workload/Workload.cpp:47-55): ranks by out-degree
order = torch.argsort(outdeg_all, descending=True, stable=True)
pick = lambda lo: order[lo: lo + max(8, a.sources)].cpu().numpy().astype(np.int32)
top, rank1k, rank1m = pick(0), pick(1000), pick(min(1_000_000, V // 4))
low = np.array([1, 2, 3, 5, 8, 13, 21, 34], dtype=np.int32)
del order
print(f"generated {need} edges on device in {tgen:.2f}s; W={wl.W} B={wl.B}; top out-degrees {outdeg_all[torch.from_numpy(top).long().cuda()].tolist()}", file=sys.stderr, flush=True)
eps = 1e-9
for kind in a.kinds.split(","):
    srcs = {"top": top, "rank1k": rank1k, "rank1m": rank1m, "low": low}[kind][: a.sources]
    t0 = time.time()
    eng = DynamicPPR(V, not a.undirected, wl.W, wl.B, srcs, epsilon=eps)
    eng.init_window_device_pairs(dev.data_ptr(), wl.W); eng.sync(); tinit = time.time() - t0
    t0 = time.time(); eng.solve_initial(); eng.sync(); tsolve = time.time() - t0
    s0 = eng.stats(0)
    t0 = time.time()
    nb_k = nb if kind != "top" else min(nb, a.top_batches)
    for k in range(nb_k):
        eng.slide_device_pairs(dev.data_ptr() + 8 * (wl.W + k * wl.B), wl.B)
    eng.sync(); wall = time.time() - t0
    rows = [eng.stats(k + 1) for k in range(nb_k)]
    f = lambda n: np.array([getattr(r, n) for r in rows], dtype=np.float64)
    ppr = f("ms_repair") + f("ms_push"); step = ppr + f("ms_window")
    T, F = f("traversed_edges"), f("frontier_pops")
    out = dict(config=5, kind=kind, sources=[int(x) for x in srcs], V=V, M=M, W=wl.W, B=wl.B, batches=nb_k, gen_s=round(tgen, 2),
               init_window_s=round(tinit, 2), initial_solve_ms=s0.ms_push, initial_traversed=int(s0.traversed_edges),
               error_flags=int(max(r.error_flags for r in rows)), wall_ms_per_batch=wall * 1e3 / nb_k,
               window_ms_mean=float(f("ms_window").mean()), repair_ms_mean=float(f("ms_repair").mean()), push_ms_mean=float(f("ms_push").mean()),
               step_ms_p50=float(np.median(step)), step_ms_p95=float(np.percentile(step, 95)),
               edge_updates_per_s_step=float(wl.B * nb_k / step.sum() * 1e3), edge_updates_per_s_ppr_only=float(wl.B * nb_k / ppr.sum() * 1e3),
               iterations=float(f("iterations").mean()), dense_sweeps=float(f("dense_sweeps").mean()), push_ms_each=[round(float(x), 2) for x in f("ms_push")], pops=float(F.mean()), traversed=float(T.mean()),
               traversed_per_update=float(T.sum() / (wl.B * nb_k * len(srcs))),
               push_edges_per_ns=float(T.sum() / max(f("ms_push").sum() * 1e6, 1e-9)),
               push_alg_GBps=float((24 * T + 56 * F).sum() / max(f("ms_push").sum() * 1e-3, 1e-12) / 1e9),
               window_alg_GBps=float(60.0 * wl.B * nb_k / (f("ms_window").sum() * 1e-3) / 1e9),
               repair_alg_GBps=float(45.0 * f("batch_entries").sum() * len(srcs) / (f("ms_repair").sum() * 1e-3) / 1e9),
               pool_used=int(rows[-1].pool_used), relocations=float(f("relocations").mean()))
    if a.check:
        t0 = time.time()
        rp, ci, od = eng.export_window_csr()
        out["csr_entries_ok"] = bool(int(rp[-1]) == wl.W and int(od.sum()) == wl.W)
        # checksum of the window against the device stream itself: sum of src*31+dst over the window, both ways
        win = dev[nb_k * wl.B: nb_k * wl.B + wl.W]
        want = int((win[:, 0].long() * 31 + win[:, 1].long()).sum().item())
        rows_ids = np.repeat(np.arange(V, dtype=np.int64), np.diff(rp))
        got = int((ci.astype(np.int64) * 31 + rows_ids).sum())
        out["window_checksum_ok"] = bool(want == got)
        p, r = eng.estimates(0), eng.residuals(0)
        out["max_abs_residual_over_eps"] = float(np.abs(r).max() / eps)
        out["invariant_defect"] = float(invariant_defect(V, rp, ci, od, p, r, int(srcs[0])))
        out["check_s"] = round(time.time() - t0, 1)
        del rp, ci, od, rows_ids, p, r
    if os.environ.get("DPPR_ITERLOG"):
        lg = eng.iterlog()
        if len(lg):
            tt = lg[:, 2].astype(np.int64); dt = np.diff(tt) / 1e3
            print(f"[{kind}] last batch per-iteration (frontier, hub_chunks|D=dense, us): " + " ".join(
                f"({int(x)},{'D' if int(y) == 0xffffffff else int(y)},{z:.0f})" for (x, y, _), z in zip(lg[:-1], dt)), file=sys.stderr)
    eng.close()
    print(json.dumps(out), flush=True)
