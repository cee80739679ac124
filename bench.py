#!/usr/bin/env python
"""bench.py -- headline benchmark of the streaming reverse-push PPR hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--sources S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one slide of the window by one batch: B edges arrive, B expire, the window graph is updated on the
device, and for EVERY source of the job residuals are repaired and both push phases run to exhaustion.

Workloads (dynamicppr_b200/workloads.py; --config picks one, default 4):
  4  BASELINE configs[3]: Orkut-shaped undirected power-law stream (3.07 M V, 117.2 M E, DRAM-resident window of
     23.4 M CSR entries), the exact top-1000 out-degree sources, -r 0.01: ONE fixed multi-source job, its source list
     split over the N GPUs (window graph replicated, no data-path collective) -> strong scaling; metric
     source_batches_per_sec (a source-batch = one source refreshed for one batch; x B = (source, edge) updates/s).
  5  BASELINE configs[4]: Twitter-shaped directed R-MAT (41.7 M V, 1.47 B E), top-64 sources, strong scaling.
  2  BASELINE configs[1]: youtube-shaped (round 1's workload, L2-resident), one source per GPU, edge_updates_per_sec.
  3  BASELINE configs[2]: LiveJournal-shaped R-MAT, mode 1 -c 100 -l 10000 (small-batch latency), --variant 0..3.

Printed JSON (one line, rank 0):
  value      whole-job throughput with the stream already resident in HBM; the clock is the sum of the per-step
             CUDA-event times on the engine's stream (window update + repair + push), max over ranks.  The working set
             exceeds L2 on configs 4 / 5 (said in config.l2); on configs 2 / 3 L2 is flushed between steps.
  e2e        the same metric through the C-ABI call a user makes (dppr_slide_pairs with HOST buffers: pinned staging +
             H2D inside the timed region) plus the read-back a query needs: the step record and the top-16 estimates
             of every source on the rank (dppr_get_topk, D2H).
  roofline   the persistent push kernel: algorithmic bytes of what it did (scatter iterations: 24 B per traversed
             in-edge + 56 B per pop; gather sweeps: 4 B per out-list entry walked + 2 B (bf16) per (entry,
             source) gather + 4 B per (vertex, source) unit (x read + x write) + per pop 16 B of r read-modify-write and 16 B of p
             (one source) or 8 B of the FP32 popped-amount sum that stands in for p during an episode (several sources))
             / CUDA-event time of that kernel, against the
             measured HBM copy bandwidth (MEASURED_PEAKS.json).  `traffic` = DRAM bytes per launch from the committed
             ncu capture of the same kernel on the same config (profiles/traffic.json), else null.
  cpu_baseline  the reference's own CPU implementation (oracle/_ref/ref_harness_omp: unmodified reference classes,
             cilk_for backed by OpenMP) on this box's host cores, same stream (.bin prefix written by the host twin of
             the device generator), same flags, a bounded sample (sources x batches stated).
`--impl reference` times that CPU implementation as the reference arm (rank 0 only): K timed + W warm-up batches for
a SAMPLE of the job's sources (--cpu-sources, quartile ranks of the source list), extrapolated to the job as
1 / mean seconds per source-batch -- the reference runs one source per process, sequentially on all cores.
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

TOPK = 16


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref), or the C restatement if it is not there
# ---------------------------------------------------------------------------------------------------------------------
def sample_sources(sources, n):
    """n sources spread over the job's list by rank (quartile-style): rank (2i+1) S / (2n)"""
    S = len(sources)
    n = max(1, min(n, S))
    return [int(sources[((2 * i + 1) * S) // (2 * n)]) for i in range(n)]


CPU_DEADLINE_S = 900.0


def run_reference_cpu(cfg, sources, n_batches, threads, drop, variant):
    """Returns per-(source, batch) times of the reference CPU classes on the same stream, or None if the binary is absent.
    The run is cut at CPU_DEADLINE_S; whatever batches it had timed by then are the sample (none: a lower bound is reported)."""
    from dynamicppr_b200 import workloads
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness_omp")
    if not os.path.exists(harness):
        return None
    wl = cfg.workload()
    d = os.path.join(tempfile.gettempdir(), "dppr_bench")
    os.makedirs(d, exist_ok=True)
    t0 = time.time()
    binfile = workloads.write_prefix_bin(cfg, os.path.join(d, f"config{cfg.index}.bin"), wl.W + (n_batches + 1) * wl.B)
    t_bin = time.time() - t0
    tfile = os.path.join(d, f"ref_times_{os.getpid()}.txt")
    cmd = [harness, "-d", binfile] + cfg.cli_flags(n_batches) + ["-s", str(sources[0]), "-t", str(threads), "-o", str(variant),
           "--quiet", "--times", tfile, "--sources", ",".join(str(s) for s in sources)]
    if not cfg.directed:
        # the reference's incremental host adjacency is wrong on undirected streams (defect D1): rebuild the true
        # window before each batch (untimed there, as the reference's own -DVALIDATE build does) so both arms work on the same graph
        cmd.append("--scratch-graph")
    t0 = time.time()
    cut = False
    try:
        subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=CPU_DEADLINE_S)
    except subprocess.TimeoutExpired:
        cut = True
    wall = time.time() - t0
    rows = [ln.split() for ln in open(tfile)] if os.path.exists(tfile) else []
    rows = [r for r in rows if len(r) == 4]
    if os.path.exists(tfile):
        os.remove(tfile)
    if not any(int(r[0]) > drop for r in rows):
        # nothing timed before the deadline: report the bound the deadline gives
        init = [float(r[1]) for r in rows if int(r[0]) == 0]
        return dict(kind="reference", s_per_source_batch=float(wall), p50_ms=float(wall * 1e3), steps=0, sources=len(sources), wall_s=wall,
                    bin_s=t_bin, init_solve_s=float(np.mean(init) / 1e6) if init else float("nan"), cut=True, bound=True)
    us = np.array([float(r[1]) for r in rows if int(r[0]) > drop])  # batch 0 = initial solve
    init_us = np.array([float(r[1]) for r in rows if int(r[0]) == 0])
    per_batch = {}
    for r in rows:
        if int(r[0]) > drop:
            per_batch.setdefault(int(r[0]), []).append(float(r[1]))
    step_ms = np.array([np.mean(v) for _, v in sorted(per_batch.items())]) / 1e3  # mean over the sampled sources, per batch
    return dict(kind="reference", s_per_source_batch=float(us.mean() / 1e6), p50_ms=float(np.median(step_ms)), steps=len(step_ms),
                sources=len(sources), wall_s=wall, bin_s=t_bin, init_solve_s=float(init_us.mean() / 1e6) if len(init_us) else float("nan"),
                cut=cut, bound=False)


def run_port_cpu(cfg, sources, n_batches, drop, variant):
    """fallback CPU baseline: the single-threaded C restatement (oracle/dppr_oracle.c)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    from dynamicppr_b200 import workloads
    wl = cfg.workload()
    edges = workloads.host_edges(cfg, 0, wl.W + (n_batches + 1) * wl.B)
    ts = []
    for s in sources[:1]:
        o = orc.Oracle(cfg.V, cfg.directed, edges, wl.W, wl.B, int(s), cfg.eps, variant)
        o.initial_solve()
        for k in range(n_batches):
            t0 = time.perf_counter(); o.slide(wl.B); dt = time.perf_counter() - t0
            if k >= drop:
                ts.append(dt)
    return dict(kind="port", s_per_source_batch=float(np.mean(ts)), p50_ms=float(np.median(ts) * 1e3), steps=len(ts), sources=1,
                wall_s=float(np.sum(ts)), bin_s=0.0, init_solve_s=0.0)


def cpu_arm(cfg, job_sources, n_cpu_sources, n_batches, drop, variant, ncores):
    picked = sample_sources(job_sources, n_cpu_sources)
    res = run_reference_cpu(cfg, picked, n_batches + drop, ncores, drop, variant)
    if res is None:
        res = run_port_cpu(cfg, picked, min(n_batches, 3) + drop, drop, variant)
    wl = cfg.workload()
    sb_per_s = 1.0 / res["s_per_source_batch"]
    sample = (f"{res['steps']} batches of {wl.B} edges after {drop} warm-up batch(es), {res['sources']} of the job's {len(job_sources)} "
              f"sources (list ranks {[int(np.where(np.asarray(job_sources) == s)[0][0]) for s in picked[:res['sources']]]}), run one after "
              f"the other with all threads each; ")
    sample += ("unmodified reference cpu/ classes (oracle/_ref/ref_harness_omp), cilk_for backed by OpenMP" if res["kind"] == "reference"
               else "single-threaded C restatement of the reference (oracle/dppr_oracle.c)")
    if res.get("bound"):
        sample += f"; NO slide finished within the {CPU_DEADLINE_S:.0f} s deadline: the value is an UPPER bound (1 / elapsed seconds)"
    elif res.get("cut"):
        sample += f"; cut at the {CPU_DEADLINE_S:.0f} s deadline, the batches timed until then are the sample"
    if res["kind"] == "reference" and not cfg.directed:
        sample += "; true window rebuilt (untimed) before each batch because the reference's incremental adjacency is wrong on undirected streams (DESIGN.md D1)"
    return res, sb_per_s, sample, picked


def main():
    # the contract is ONE JSON line on stdout: libraries that chat on fd 1 (NCCL prints its version there) are
    # diverted to stderr, the JSON goes to the saved descriptor
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5])
    ap.add_argument("--sources", type=int, default=0, help="sources of the whole job (default: the config's: 1000 / 64; configs 2, 3: per GPU)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-flush", action="store_true", help="configs 2 / 3: do not flush L2 between steps")
    ap.add_argument("--cpu-sources", type=int, default=2, help="sources of the job the CPU arm samples")
    ap.add_argument("--cpu-batches", type=int, default=3, help="timed batches of the in-line cpu_baseline (the reference arm uses --steps)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end loop (default: --steps, or 5 when a step exceeds 0.2 s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-deadline", type=float, default=CPU_DEADLINE_S, help="seconds after which the CPU arm is cut")
    a = ap.parse_args()
    globals()["CPU_DEADLINE_S"] = a.cpu_deadline
    a.warmup = max(a.warmup, 3)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1

    from dynamicppr_b200 import workloads
    cfg = workloads.CONFIGS[a.config]
    wl = cfg.workload()
    multi = cfg.multi_source
    if multi:
        S_total = a.sources if a.sources > 0 else cfg.n_sources
        metric, unit, scaling = "source_batches_per_sec", "source-batches/s", "strong"
    else:
        spg = a.sources if a.sources > 0 else 1
        S_total = spg * world
        metric, unit, scaling = "edge_updates_per_sec", "edge updates/s", "weak"
    avail = wl.runnable_batches(cfg.M)
    resident = cfg.index in (2, 3)  # working set fits the 126 MB L2: flush between steps
    config = {"workload": cfg.describe() + f", {'top-' + str(S_total) + ' out-degree sources' if multi else 'top out-degree source(s)'}, -o {a.variant}",
              "variant": a.variant, "sources_total": S_total, "V": cfg.V, "M": cfg.M, "W": wl.W, "B": wl.B,
              "parallelism": (f"{S_total} sources dealt round-robin (by degree rank) to {world} GPU(s), window graph replicated, no data-path collective" if multi
                              else f"source-sharded x{world} ({S_total // world} per GPU), window graph replicated"),
              "l2": ("flushed between steps (256 MiB write, outside the timed events)" if resident and not a.no_flush else
                     "not flushed: stateful stream" if resident else
                     f"not flushed: per-GPU working set ({16 * cfg.V * max(S_total // world, 1) / 1e9:.1f} GB of p, r + the window graph) exceeds the 126 MB L2")}

    # ------------------------------------------------------------------ reference arm (CPU)
    if a.impl == "reference":
        if rank != 0:
            return
        job_sources = workloads.top_sources(cfg, S_total, on_host=not _cuda_ok())
        steps = min(a.steps, avail - a.warmup)
        res, sb_per_s, sample, picked = cpu_arm(cfg, job_sources, a.cpu_sources if multi else 1, steps, a.warmup, a.variant, ncores)
        value = sb_per_s if multi else sb_per_s * wl.B
        cores = ncores if res["kind"] == "reference" else 1
        config["reference_sampling"] = (f"value = 1 / mean seconds per source-batch over {res['sources']} sampled source(s) x {res['steps']} batches"
                                        + (" (x B edges)" if not multi else "") + "; the reference processes a job's sources one process at a time")
        out = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": a.gpus, "steps": res["steps"],
               "warmup": a.warmup, "ms_per_step": res["s_per_source_batch"] * 1e3 * (S_total if multi else 1), "p50_ms": res["p50_ms"],
               "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": res["kind"], "sample": sample,
                                "s_per_source_batch": res["s_per_source_batch"], "init_solve_s": res["init_solve_s"]},
               "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    t0 = time.time()
    job_sources = workloads.top_sources(cfg, S_total, device=local_rank)
    my_sources = sharding.shard_sources(job_sources, rank, world, None if multi else S_total // world, interleave=multi)
    K, Wm = a.steps, a.warmup
    if Wm + 2 * K > avail:
        K = max(1, (avail - Wm) // 2)
    n_dev = wl.W + (Wm + K) * wl.B
    dev_edges = workloads.device_edges(cfg, 0, n_dev, local_rank)
    torch.cuda.synchronize()
    log(f"config {cfg.index}: V={cfg.V} M={cfg.M} W={wl.W} B={wl.B}; {len(my_sources)} of {S_total} sources on this rank; "
        f"stream + source ranking ready in {time.time() - t0:.1f}s")
    t0 = time.time()
    eng = DynamicPPR(cfg.V, cfg.directed, wl.W, wl.B, my_sources, epsilon=cfg.eps, variant=a.variant, device=local_rank, record_timing=True)
    eng.init_window_device_pairs(dev_edges.data_ptr(), wl.W)
    eng.solve_initial()
    eng.sync()
    s0 = eng.stats(0)
    log(f"engine ready in {time.time() - t0:.1f}s (initial solve {s0.ms_push:.1f} ms, {s0.iterations} iterations)")
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if resident and not a.no_flush else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush():
        if flush_buf is not None:
            flush_buf.zero_()
            torch.cuda.synchronize()

    def dev_step(k):
        eng.slide_device_pairs(dev_edges.data_ptr() + 8 * (wl.W + k * wl.B), wl.B)

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

    # ---- end to end: host buffers through the public C-ABI call; the step record and the top-k estimates of every
    # source of this rank are read back inside the timed region
    Ke = a.e2e_steps if a.e2e_steps > 0 else (K if dev_total_ms / K < 200.0 else min(K, 5))
    Ke = max(1, min(Ke, avail - Wm - K))
    host_edges = np.ascontiguousarray(workloads.host_edges(cfg, wl.W + (Wm + K) * wl.B, Ke * wl.B))
    e2e_t = []
    for k in range(Ke):
        flush()
        t0 = time.perf_counter()
        eng.slide_pairs(host_edges[k * wl.B:(k + 1) * wl.B])
        st = eng.stats()  # synchronises and reads the batch record back from pinned host memory
        ids, vals = eng.topk(TOPK)
        e2e_t.append(time.perf_counter() - t0)
        if st.error_flags:
            raise SystemExit(f"device error flags {st.error_flags}: results invalid")
    e2e_ms_per_step = float(np.mean(e2e_t)) * 1e3
    log("e2e ms per step: " + " ".join(f"{t * 1e3:.2f}" for t in e2e_t[:12]))
    assert ids.shape == (len(my_sources), TOPK) and np.all(vals[:, 0] > 0)

    # ---- max over ranks; the job's only data collective (after the timed region): top-k digests of every source and
    # one full estimate vector per rank to rank 0
    gathered_rows = len(my_sources)
    if world > 1:
        dev_total_ms, e2e_ms_per_step, t_wall = sharding.max_over_ranks([dev_total_ms, e2e_ms_per_step, t_wall], device="cuda")
        digest = torch.from_numpy(np.concatenate([ids.astype(np.float64), vals], axis=1)).cuda()
        got = sharding.gather_estimates(digest, dst=0)
        full = torch.empty((1, cfg.V), dtype=torch.float64, device="cuda")
        eng.copy_estimates_device(0, full.data_ptr())
        fulls = sharding.gather_estimates(full, dst=0)
        if rank == 0:
            gathered_rows = sum(int(g.shape[0]) for g in got)
            assert len(fulls) == world
    n_units = float(S_total * K) if multi else float(S_total * K * wl.B)
    value = n_units / (dev_total_ms * 1e-3)
    e2e_value = (n_units / K) / (e2e_ms_per_step * 1e-3)

    if rank == 0:
        T, F = f("traversed_edges"), f("frontier_pops")
        Ts, Fd = f("scatter_edges"), f("dense_pops")
        L, Ls, U = f("dense_slots"), f("dense_pairs"), f("dense_units")
        push_s = float(f("ms_push").sum()) * 1e-3
        # sources beyond 128 are refreshed in equal panels, one push launch each (engine.cuh, Pw_): the roofline is per launch
        panel_max = int(os.environ.get("DPPR_PANEL_SOURCES", "128") or 128)
        n_panels = -(-len(my_sources) // panel_max)
        panel_w = -(-len(my_sources) // n_panels)
        pop_bytes = 32.0 if panel_w == 1 else 24.0  # (a sweep's pop: r RMW + p RMW, or r RMW + the FP32 amount sum, pull.cuh)
        alg_bytes = float((24.0 * Ts + 56.0 * (F - Fd) + 4.0 * L + 2.0 * Ls + 4.0 * U + pop_bytes * Fd).sum())
        peak, peak_src = measured_hbm_peak()
        achieved = alg_bytes / push_s / 1e9
        dense = bool(f("dense_sweeps").sum() > 0)
        traffic, traffic_note = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get(f"config{cfg.index}_s{panel_w}_o{a.variant}")
            if ent:
                traffic, traffic_note = ent.get("dram_bytes_per_launch"), ent.get("source")
        except Exception:
            pass
        # which instantiation ran (engine.cu, Engine::Engine): variants 1-3 always get the kernel that can switch to sweeps,
        # variant 0 from 2e7 window entries x sources per launch
        switching = a.variant != 0 or wl.W * (1 if cfg.directed else 2) * panel_w >= 2.0e7
        kname = f"push_persistent<{a.variant}, {(1 if panel_w == 1 else 8) if switching else 0}>"
        roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes / K / n_panels, "launch_ms": push_s * 1e3 / K / n_panels,
                    "launches_per_step": n_panels, "sources_per_launch": panel_w,
                    "scatter_form_equivalent_GBps": float((24.0 * T + 56.0 * F).sum()) / push_s / 1e9,
                    "note": (f"bytes = 24 T_scatter + 56 F_scatter + 4 slots + 2 (slot, source) bf16 gathers + 4 (vertex, source) units + {pop_bytes:.0f} F_dense "
                             "(DESIGN.md 3.3); 'scatter_form_equivalent' credits every gathered non-zero pair the 24 B a scatter would move "
                             "and is NOT the roofline figure" if dense else
                             "scatter iterations only: 24 B per traversed in-edge + 56 B per pop; L2-resident working set: bound by dependent "
                             "L2 round trips and FP64 atomics, not DRAM (DESIGN.md 3.3)")}
        if not dense:
            # scatter levels are one returning FP64 atomic per traversed in-edge: measured ceiling on random L2-resident
            # addresses 117 / ns (profiles/micro_fp64_atomics_r01.txt) -- the bound that applies when the working set fits L2
            try:
                per_ns = float(T.sum()) / (push_s * 1e9)
                roofline.update({"atomics_per_ns": per_ns, "atomic_ceiling_per_ns": 117.0, "frac_of_atomic_ceiling": per_ns / 117.0})
            except Exception:
                pass
        cpu = None
        if world == 1 and not a.no_cpu:
            res, sb_per_s, sample, _ = cpu_arm(cfg, job_sources, a.cpu_sources if multi else 1, a.cpu_batches, 1, a.variant, ncores)
            cpu = {"value": sb_per_s if multi else sb_per_s * wl.B, "unit": unit, "cores": ncores if res["kind"] == "reference" else 1,
                   "kind": res["kind"], "s_per_source_batch": res["s_per_source_batch"], "init_solve_s": res["init_solve_s"], "sample": sample}
        out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": Wm,
               "ms_per_step": dev_total_ms / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": 8 * wl.B,
                       "d2h_bytes_per_step": 160 + len(my_sources) * TOPK * 12, "ms_per_step": e2e_ms_per_step, "steps": Ke,
                       "readback": f"step record + top-{TOPK} (id, estimate) of each of the rank's {len(my_sources)} sources"},
               "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
               "p50_ms": float(np.median(step_ms)), "p95_ms": float(np.percentile(step_ms, 95)),
               "edge_updates_per_sec": float(wl.B * K / (dev_total_ms * 1e-3)),
               "source_edge_updates_per_sec": float(S_total * wl.B * K / (dev_total_ms * 1e-3)),
               "ppr_only": {"ms_per_step": float(ppr_ms.mean()), "p50_ms": float(np.median(ppr_ms)),
                            "note": "the reference's own clock: repair + push, window update excluded (gpu/PPRGPU.cuh:128-163)"},
               "per_step": {"iterations": float(f("iterations").mean()), "dense_sweeps": float(f("dense_sweeps").mean()),
                            "frontier_pops": float(F.mean()), "traversed_edges": float(T.mean()), "scatter_edges": float(Ts.mean()),
                            "ms_window": float(f("ms_window").mean()), "ms_repair": float(f("ms_repair").mean()),
                            "ms_push": float(f("ms_push").mean()), "relocations": float(f("relocations").mean()),
                            "pool_used": int(rows[-1].pool_used)},
               "initial_solve_ms": float(s0.ms_push), "wall_ms_per_step": t_wall * 1e3 / K, "estimates_gathered": gathered_rows}
        print(json.dumps(out), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _cuda_ok():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


if __name__ == "__main__":
    main()
