import sys, os, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from dynamicppr_b200 import DynamicPPR, DpprError
from test_gpu_round2 import _drifting_hub_stream
V, W, B = 4_000, 40_000, 2_000
rng = np.random.default_rng(5)
M = W + 22 * W
edges = _drifting_hub_stream(V, M, W // 2, rng)
nb = (M - W) // B
eng = DynamicPPR(V, True, W, B, [1], epsilon=1e-7, pool_factor=2.0, tuning={"dense": 1, "iterlog": 1})
eng.init_window_pairs(edges[:W]); eng.solve_initial()
for k in range(nb):
    try:
        eng.slide_pairs(edges[W + k * B: W + (k + 1) * B]); eng.sync()
    except DpprError as e:
        print("FAILED at batch", k + 1, e)
        st = eng.stats()
        print(st.as_dict())
        lg = eng.iterlog()
        print("iterlog rows", len(lg)); print(lg[:40]); print(lg[-10:])
        break
    st = eng.stats()
    if k % 50 == 0: print(k, st.iterations, st.dense_sweeps, st.ms_push)
else:
    print("all ok")
