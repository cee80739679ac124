#!/usr/bin/env python
"""Run one BASELINE.json config (dynamicppr_b200/workloads.py) on one GPU with the stream generated on the device, and
print one JSON line: timings, counters, the device-side parity checks, optionally the per-iteration log of the last batch.

    python scripts/run_baseline.py --config 5 --sources 1 --batches 3 [--rank-offset 1000] [--iterlog] [--variant 0]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dynamicppr_b200 import DynamicPPR, workloads

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True); ap.add_argument("--sources", type=int, default=1)
ap.add_argument("--rank-offset", type=int, default=0, help="first degree rank of the source list (0 = the top sources)")
ap.add_argument("--batches", type=int, default=5); ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--check", type=int, default=1); ap.add_argument("--iterlog", action="store_true")
ap.add_argument("--tuning", default="", help="comma-separated dppr_tuning fields, e.g. dense=-1,tile_cap=64")
a = ap.parse_args()
cfg = workloads.CONFIGS[a.config]
wl = cfg.workload()
nb = min(a.batches, wl.runnable_batches(cfg.M))
t0 = time.time()
order = workloads.top_sources(cfg, a.rank_offset + a.sources)
srcs = order[a.rank_offset:]
dev = workloads.device_edges(cfg, 0, wl.W + nb * wl.B)
torch.cuda.synchronize(); tgen = time.time() - t0
tuning = {"iterlog": 1} if a.iterlog else {}
for kv in filter(None, a.tuning.split(",")):
    k, v = kv.split("="); tuning[k] = float(v) if "." in v or "e" in v else int(v)
t0 = time.time()
eng = DynamicPPR(cfg.V, cfg.directed, wl.W, wl.B, srcs, epsilon=cfg.eps, variant=a.variant, tuning=tuning)
eng.init_window_device_pairs(dev.data_ptr(), wl.W); eng.sync(); tinit = time.time() - t0
eng.solve_initial(); eng.sync(); s0 = eng.stats(0)
t0 = time.time()
for k in range(nb):
    eng.slide_device_pairs(dev.data_ptr() + 8 * (wl.W + k * wl.B), wl.B)
eng.sync(); wall = time.time() - t0
rows = [eng.stats(k + 1) for k in range(nb)]
f = lambda n: np.array([getattr(r, n) for r in rows], dtype=np.float64)
T, F = f("traversed_edges"), f("frontier_pops")
alg = 24 * f("scatter_edges") + 56 * (F - f("dense_pops")) + 4 * f("dense_slots") + 2 * f("dense_pairs") + 4 * f("dense_units") + (32 if len(srcs) == 1 else 24) * f("dense_pops")
out = dict(config=a.config, sources=len(srcs), rank_offset=a.rank_offset, variant=a.variant, tuning=tuning, V=cfg.V, W=wl.W, B=wl.B, batches=nb,
           gen_s=round(tgen, 2), init_window_s=round(tinit, 2), initial_solve_ms=s0.ms_push, initial_iterations=int(s0.iterations),
           error_flags=int(max(r.error_flags for r in rows)), wall_ms_per_batch=wall * 1e3 / nb,
           window_ms=float(f("ms_window").mean()), repair_ms=float(f("ms_repair").mean()), push_ms=float(f("ms_push").mean()),
           push_ms_each=[round(float(x), 2) for x in f("ms_push")],
           edge_updates_per_s=float(wl.B * nb / (f("ms_window") + f("ms_repair") + f("ms_push")).sum() * 1e3),
           source_batches_per_s=float(len(srcs) * nb / (f("ms_window") + f("ms_repair") + f("ms_push")).sum() * 1e3),
           iterations=float(f("iterations").mean()), dense_sweeps=float(f("dense_sweeps").mean()), pops=float(F.mean()),
           traversed=float(T.mean()), scatter_edges=float(f("scatter_edges").mean()), dense_pairs=float(f("dense_pairs").mean()),
           push_alg_GBps=float(alg.sum() / (f("ms_push").sum() * 1e-3) / 1e9),
           repair_alg_GBps=float(45.0 * f("batch_entries").sum() * len(srcs) / (f("ms_repair").sum() * 1e-3) / 1e9),
           window_alg_GBps=float(60.0 * (1 if cfg.directed else 2) * wl.B * nb / (f("ms_window").sum() * 1e-3) / 1e9),
           pool_used=int(rows[-1].pool_used), relocations=float(f("relocations").mean()))
if a.check:
    t0 = time.time()
    out["window_mismatches"] = eng.check_window_device(dev.data_ptr() + 8 * nb * wl.B, wl.W)
    v = [eng.validate(int(i)) for i in sorted(set(np.linspace(0, len(srcs) - 1, 6).astype(int).tolist()))]  # (spread over the source panels)
    out["max_abs_residual_over_eps"] = max(x[0] for x in v) / cfg.eps
    out["invariant_defect"] = max(x[1] for x in v)
    out["check_s"] = round(time.time() - t0, 2)
if a.iterlog:
    lg = eng.iterlog()
    if len(lg):
        tt = lg[:, 2].astype(np.int64); dt = np.diff(tt) / 1e3
        kinds = ["D" if int(y) == 0xffffffff else "S" for y in lg[:-1, 1]]
        d_us = sum(z for k, z in zip(kinds, dt) if k == "D"); s_us = sum(z for k, z in zip(kinds, dt) if k == "S")
        out["last_batch"] = dict(dense_iters=kinds.count("D"), dense_us=round(d_us), scatter_iters=kinds.count("S"), scatter_us=round(s_us))
        print("last batch per-iteration (frontier, hub_chunks|D, us): " + " ".join(
            f"({int(x)},{'D' if int(y) == 0xffffffff else int(y)},{z:.0f})" for (x, y, _), z in zip(lg[:-1], dt)), file=sys.stderr)
eng.close()
print(json.dumps(out), flush=True)
