#!/usr/bin/env python
"""bench.py -- headline benchmark of the streaming reverse-push PPR hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one slide of the window by one batch: B edges arrive, B expire, the window graph is
updated on the device, residuals are repaired and both push phases run to exhaustion.
Workload (config.workload): BASELINE.json configs[1] -- youtube-shaped synthetic undirected
power-law stream (1,134,890 V, 2,987,624 E), window 0.1, -r 0.01 (B = 2,987 edges), eps 1e-9,
top-out-degree source, variant -o 0.  One process per GPU; with N > 1 every rank replicates the
window graph and owns its own source(s) (weak scaling, no collective on the data path; the final
gather of estimates over NCCL happens after the timed region).

Printed JSON (one line, rank 0):
  value      whole-job edge updates/s with the stream already resident in HBM; the clock is the sum
             of the per-step CUDA-event times on the engine's stream (window update + repair +
             push), max over ranks.  L2 is flushed between steps (outside the events).
  e2e        same metric through the C-ABI call a user makes (dppr_slide_pairs with HOST buffers:
             pinned staging + H2D inside the timed region, plus a D2H read of the step's record).
  roofline   push kernel (push_persistent<0, false>): algorithmic bytes (24 B per traversed in-edge +
             56 B per frontier pop) / CUDA-event time of that kernel, vs the measured
             HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's own CPU implementation (oracle/_ref/ref_harness_omp: unmodified
             reference sources, cilk_for backed by OpenMP) on this box's host cores, same stream,
             same flags, bounded number of batches.
`--impl reference` times that CPU implementation as the reference arm (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = "youtube"
WINDOW_RATIO, BATCH_RATIO, EPS = 0.1, 0.01, 1e-9
METRIC, UNIT = "edge_updates_per_sec", "edge updates/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(nsrc_total):
    from dynamicppr_b200 import graphgen, stream
    V, M, directed = graphgen.SHAPES[SHAPE]
    seed = graphgen.BASE_SEED + list(graphgen.SHAPES).index(SHAPE)
    t0 = time.time()
    edges = graphgen.powerlaw_undirected(V, M, seed)
    wl = stream.workload(M, WINDOW_RATIO, 0, BATCH_RATIO, 10 ** 9)
    sources = graphgen.top_out_degree(V, edges, directed, max(nsrc_total, 1))
    log(f"[bench] {SHAPE}-shaped stream V={V} M={M} W={wl.W} B={wl.B} generated in {time.time() - t0:.1f}s")
    return V, M, directed, edges, wl, sources


def bin_path(V, edges):
    from dynamicppr_b200 import graphgen
    d = os.path.join(tempfile.gettempdir(), "dppr_bench")
    os.makedirs(d, exist_ok=True)
    p = os.path.join(d, f"{SHAPE}.bin")
    if not os.path.exists(p) or os.path.getsize(p) != 4 + 8 * len(edges):
        graphgen.write_bin(p, V, edges)
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference_cpu(binfile, source, n_batches, threads, drop):
    """the reference's CPU implementation of the path (oracle/_ref, built from the unmodified sources)"""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness_omp")
    kind = "reference"
    if not os.path.exists(harness):
        return None
    tfile = os.path.join(tempfile.gettempdir(), f"dppr_ref_times_{os.getpid()}.txt")
    cmd = [harness, "-d", binfile, "-a", "0", "-i", "0", "-y", "1", "-n", "0", "-w", str(WINDOW_RATIO), "-r",
           str(BATCH_RATIO), "-b", str(n_batches), "-s", str(source), "-t", str(threads), "-o", "0", "-e", repr(EPS),
           "--quiet", "--times", tfile]
    t0 = time.time()
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    wall = time.time() - t0
    rows = [ln.split() for ln in open(tfile)]
    os.remove(tfile)
    us = np.array([float(r[1]) for r in rows[1:]])  # row 0 = initial solve
    us = us[drop:] if len(us) > drop else us
    return dict(kind=kind, ms_per_step=float(us.mean() / 1e3), p50_ms=float(np.median(us) / 1e3), steps=len(us), wall_s=wall)


def run_port_cpu(V, directed, edges, wl, source, n_batches):
    """fallback CPU baseline: the single-threaded C restatement (oracle/dppr_oracle.c)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    o = orc.Oracle(V, directed, edges, wl.W, wl.B, source, EPS, 0)
    o.initial_solve()
    ts = []
    for _ in range(n_batches):
        t0 = time.perf_counter(); o.slide(wl.B); ts.append(time.perf_counter() - t0)
    return dict(kind="port", ms_per_step=float(np.mean(ts) * 1e3), p50_ms=float(np.median(ts) * 1e3), steps=n_batches, wall_s=sum(ts))


def main():
    # the contract is ONE JSON line on stdout: libraries that chat on fd 1 (NCCL prints its version there) are
    # diverted to stderr, the JSON goes to the saved descriptor
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sources-per-gpu", type=int, default=1)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between steps")
    ap.add_argument("--cpu-batches", type=int, default=40, help="batches of the bounded CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1

    config = {"workload": f"BASELINE configs[1]: {SHAPE}-shaped synthetic undirected power-law stream, window {WINDOW_RATIO}, "
                          f"-r {BATCH_RATIO} (mode 0), eps {EPS}, top-out-degree source(s), -o {a.variant}",
              "variant": a.variant, "sources_per_gpu": a.sources_per_gpu, "parallelism": f"source-sharded x{world}, window graph replicated",
              "l2": "flushed between steps (256 MiB write, outside the timed events)" if not a.no_flush else "not flushed (stateful stream)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        V, M, directed, edges, wl, sources = make_workload(1)
        config.update(V=V, M=M, W=wl.W, B=wl.B)
        binfile = bin_path(V, edges)
        steps = min(a.steps, wl.runnable_batches(M) - a.warmup)
        res = run_reference_cpu(binfile, int(sources[0]), steps + a.warmup, ncores, a.warmup)
        if res is None:
            res = run_port_cpu(V, directed, edges, wl, int(sources[0]), min(steps, 5))
        value = wl.B / (res["ms_per_step"] * 1e-3)
        cores = ncores if res["kind"] == "reference" else 1
        out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": res["steps"],
               "warmup": a.warmup, "ms_per_step": res["ms_per_step"], "p50_ms": res["p50_ms"], "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": res["kind"],
                                "sample": f"{res['steps']} batches of {wl.B} edges after {a.warmup} warm-up batches, 1 source; "
                                          "reference cpu/ sources, cilk_for backed by OpenMP" if res["kind"] == "reference"
                                          else f"{res['steps']} batches, single-threaded C restatement"},
               "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out), file=real_stdout, flush=True)
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    import torch.distributed as dist
    from dynamicppr_b200 import DynamicPPR, binding, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    spg = a.sources_per_gpu
    V, M, directed, edges, wl, sources = make_workload(world * spg)
    config.update(V=V, M=M, W=wl.W, B=wl.B)
    my_sources = sharding.shard_sources(sources, rank, world, spg)
    avail = wl.runnable_batches(M)
    K, Wm = a.steps, a.warmup
    if Wm + 2 * K > avail:
        K = (avail - Wm) // 2
    eng = DynamicPPR(V, directed, wl.W, wl.B, my_sources, epsilon=EPS, variant=a.variant, device=local_rank, record_timing=True)
    eng.init_window_pairs(edges[: wl.W])
    eng.solve_initial()
    eng.sync()

    dev_edges = torch.from_numpy(np.ascontiguousarray(edges[wl.W: wl.W + (Wm + K) * wl.B])).cuda()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush():
        if not a.no_flush:
            flush_buf.zero_()
            torch.cuda.synchronize()

    def dev_step(k):
        eng.slide_device_pairs(dev_edges.data_ptr() + 8 * k * wl.B, wl.B)

    for k in range(Wm):  # warm-up (untimed)
        flush(); dev_step(k)
    eng.sync()
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    launches0 = binding.kernel_launches()
    t_wall0 = time.perf_counter()
    for k in range(Wm, Wm + K):
        flush(); dev_step(k); eng.sync()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = binding.kernel_launches() - launches0
    rows = [eng.stats(k + 1) for k in range(Wm, Wm + K)]
    clocks = sampler.stop()
    err = max(r.error_flags for r in rows)
    if err:
        raise SystemExit(f"device error flags {err}: results invalid")
    f = lambda name: np.array([getattr(r, name) for r in rows], dtype=np.float64)
    step_ms = f("ms_window") + f("ms_repair") + f("ms_push")
    ppr_ms = f("ms_repair") + f("ms_push")
    dev_total_ms = float(step_ms.sum())

    # ---- end to end: host buffers through the public C-ABI call, D2H of the step record inside the timed region
    host_edges = np.ascontiguousarray(edges[wl.W + (Wm + K) * wl.B: wl.W + (Wm + 2 * K) * wl.B])
    e2e_t = []
    for k in range(K):
        flush()
        t0 = time.perf_counter()
        eng.slide_pairs(host_edges[k * wl.B:(k + 1) * wl.B])
        st = eng.stats()  # synchronises and reads the batch record back from pinned host memory
        e2e_t.append(time.perf_counter() - t0)
        if st.error_flags:
            raise SystemExit(f"device error flags {st.error_flags}: results invalid")
    e2e_total_s = float(np.sum(e2e_t))

    # ---- max over ranks
    gathered_rows = spg
    if world > 1:
        dev_total_ms, e2e_total_s, t_wall = sharding.max_over_ranks([dev_total_ms, e2e_total_s, t_wall], device="cuda")
        # the only data collective of the job: gather the estimate vectors (after the timed region)
        mine = torch.empty((spg, V), dtype=torch.float64, device="cuda")
        for i in range(spg):
            eng.copy_estimates_device(i, mine[i].data_ptr())
        gathered = sharding.gather_estimates(mine, dst=0)
        if rank == 0:
            gathered_rows = sum(int(g.shape[0]) for g in gathered)
    units = float(world * spg * K * wl.B)  # (source, edge) updates processed by the whole job
    value = units / (dev_total_ms * 1e-3)
    e2e_value = units / e2e_total_s

    if rank == 0:
        T, F = f("traversed_edges"), f("frontier_pops")
        push_s = float(f("ms_push").sum()) * 1e-3
        alg_bytes = float((24.0 * T + 56.0 * F).sum())
        peak, peak_src = measured_hbm_peak()
        achieved = alg_bytes / push_s / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_push.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {"bound": "hbm", "kernel": f"push_persistent<{a.variant}, false>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes / K, "launch_ms": push_s * 1e3 / K,
                    "note": "working set (p, r, window graph ~50 MB) is L2-resident on B200: the kernel is bound by dependent "
                            "L2 round trips and FP64 atomics, not by DRAM; see DESIGN.md"}
        cpu = None
        if world == 1 and not a.no_cpu:
            binfile = bin_path(V, edges)
            res = run_reference_cpu(binfile, int(my_sources[0]), a.cpu_batches + 2, ncores, 2)
            if res is None:
                res = run_port_cpu(V, directed, edges, wl, int(my_sources[0]), 3)
            cpu = {"value": wl.B / (res["ms_per_step"] * 1e-3), "unit": UNIT, "cores": ncores if res["kind"] == "reference" else 1,
                   "kind": res["kind"], "ms_per_step": res["ms_per_step"],
                   "sample": f"{res['steps']} batches of {wl.B} edges of the same stream, same source, after 2 warm-up batches"
                             + ("; unmodified reference cpu/ sources with cilk_for backed by OpenMP" if res["kind"] == "reference" else
                                "; single-threaded C restatement of the reference")}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
               "ms_per_step": dev_total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * wl.B, "d2h_bytes_per_step": 128,
                       "ms_per_step": e2e_total_s * 1e3 / K},
               "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
               "p50_ms": float(np.median(step_ms)), "p95_ms": float(np.percentile(step_ms, 95)),
               "ppr_only": {"value": units / (float(ppr_ms.sum()) * 1e-3), "unit": UNIT, "p50_ms": float(np.median(ppr_ms)),
                            "note": "the reference's own clock: repair + push, window update excluded (gpu/PPRGPU.cuh:128-163)"},
               "per_step": {"iterations": float(f("iterations").mean()), "frontier_pops": float(F.mean()),
                            "traversed_edges": float(T.mean()), "ms_window": float(f("ms_window").mean()),
                            "ms_repair": float(f("ms_repair").mean()), "ms_push": float(f("ms_push").mean())},
               "wall_ms_per_step_incl_flush": t_wall * 1e3 / K, "estimates_gathered": gathered_rows}
        print(json.dumps(out), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
