import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import torch
from dynamicppr_b200 import DynamicPPR, workloads
cfg = workloads.CONFIGS[4]; wl = cfg.workload()
srcs = workloads.top_sources(cfg, 125)
dev = workloads.device_edges(cfg, 0, wl.W + 2 * wl.B)
eng = DynamicPPR(cfg.V, cfg.directed, wl.W, wl.B, srcs, epsilon=cfg.eps, tuning={"iterlog": 1, "probe_iter": 100000})
eng.init_window_device_pairs(dev.data_ptr(), wl.W); eng.solve_initial(); eng.sync()
cl0 = eng.ctalog().astype(np.float64)
eng.slide_device_pairs(dev.data_ptr() + 8 * wl.W, wl.B); eng.sync()
cl = eng.ctalog().astype(np.float64) - cl0
st = eng.stats()
print("push ms", st.ms_push, "sweeps", st.dense_sweeps, "CTAs", len(cl))
tot = cl.sum(0)
print("items per CTA mean", cl[:,0].mean(), "edges per item", tot[5]/tot[0])
for i, nm in enumerate(["setup", "staging", "gather", "finish"]):
    print(f"{nm:8s} mean per item {tot[1+i]/tot[0]/1e3:7.2f} us   per CTA total {cl[:,1+i].mean()/1e6:8.2f} ms")
print("sum of phases per CTA (ms): mean", cl[:,1:5].sum(1).mean()/1e6, "max", cl[:,1:5].sum(1).max()/1e6)
