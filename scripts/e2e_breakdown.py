#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its wall time: host call, device step, record read-back, top-k.
    python scripts/e2e_breakdown.py --config 4 --sources 125"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamicppr_b200 import DynamicPPR, workloads

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4); ap.add_argument("--sources", type=int, default=125)
ap.add_argument("--steps", type=int, default=6); ap.add_argument("--k", type=int, default=16)
a = ap.parse_args()
cfg = workloads.CONFIGS[a.config]; wl = cfg.workload()
srcs = workloads.top_sources(cfg, a.sources, device=0)
dev = workloads.device_edges(cfg, 0, wl.W, 0)
host = np.ascontiguousarray(workloads.host_edges(cfg, wl.W, a.steps * wl.B))
eng = DynamicPPR(cfg.V, cfg.directed, wl.W, wl.B, srcs, epsilon=cfg.eps, record_timing=True)
eng.init_window_device_pairs(dev.data_ptr(), wl.W); eng.solve_initial(); eng.sync()
print("step | slide_pairs call | stats() wait | device step | topk | total (ms)")
for k in range(a.steps):
    t0 = time.perf_counter(); eng.slide_pairs(host[k * wl.B:(k + 1) * wl.B])
    t1 = time.perf_counter(); st = eng.stats()
    t2 = time.perf_counter(); ids, vals = eng.topk(a.k)
    t3 = time.perf_counter()
    print(f"{k:4d} | {1e3 * (t1 - t0):8.3f} | {1e3 * (t2 - t1):8.3f} | {st.ms_window + st.ms_repair + st.ms_push:8.3f} | {1e3 * (t3 - t2):8.3f} | {1e3 * (t3 - t0):8.3f}")
