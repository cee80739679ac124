"""GPU: the `pagerank` CLI end to end (same flags and stdout keys as the reference's gpu/pagerank)."""
import os
import subprocess

import numpy as np
import pytest

import oracle as orc
from dynamicppr_b200 import graphgen, stream

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "dynamicppr_b200", "bin", "pagerank")


def _kv(out):
    kv = {}
    for line in out.splitlines():
        t = line.split()
        if len(t) == 2:
            kv[t[0]] = t[1]
    return kv


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_cli_dynamic_run_matches_oracle(tmp_path, variant):
    V, M = 6000, 50000
    e = graphgen.powerlaw_undirected(V, M, seed=31)
    p = str(tmp_path / "g.bin"); dump = str(tmp_path / "p.f64")
    graphgen.write_bin(p, V, e)
    src = int(graphgen.top_out_degree(V, e, False, 1)[0])
    r = subprocess.run([CLI, "-d", p, "-a", "0", "-i", "0", "-y", "1", "-w", "0.1", "-n", "0", "-r", "0.02", "-b", "12",
                        "-s", str(src), "-o", str(variant), "-e", "1e-9", "--dump", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:]
    kv = _kv(r.stdout)
    wl = stream.workload(M, 0.1, 0, 0.02, 12)
    assert int(kv["edge_num"]) == wl.B * 12 and int(kv["batches_done"]) == 12
    assert float(kv["ppr_latency"]) > 0 and float(kv["ppr_throughput"]) > 0
    assert abs(float(kv["ppr_time"]) / 12 - float(kv["ppr_latency"])) < 1e-3 * float(kv["ppr_latency"]) + 1e-9
    lines = r.stdout.strip().splitlines()
    assert lines[-1].startswith("ppr_throughput ") and lines[-2].startswith("ppr_latency ")  # scraped keys come last
    o = orc.Oracle(V, False, e, wl.W, wl.B, src, 1e-9, 0)
    o.initial_solve()
    for _ in range(12):
        o.slide(wl.B)
    got = np.fromfile(dump, dtype=np.float64)
    assert got.shape == (V,)
    assert np.abs(got - o.p).max() <= 2e-9


def test_cli_stops_when_the_stream_runs_out_and_static_mode(tmp_path):
    V, M = 500, 2000
    e = graphgen.powerlaw_undirected(V, M, seed=32)
    p = str(tmp_path / "g.bin")
    graphgen.write_bin(p, V, e)
    r = subprocess.run([CLI, "-d", p, "-i", "0", "-y", "1", "-w", "0.5", "-n", "0", "-r", "0.3", "-b", "10", "-s", "1"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:]
    assert int(_kv(r.stdout)["batches_done"]) == 3  # (M - W) // B, SlidingGraphVec.h:221
    # -y 0: from-scratch solve on the whole file (the reference GPU build cannot do this: PPRGPU.cuh:31-32)
    dump = str(tmp_path / "s.f64")
    r = subprocess.run([CLI, "-d", p, "-i", "0", "-y", "0", "-r", "0.1", "-b", "1", "-s", "3", "--dump", dump],
                       capture_output=True, text=True)
    assert r.returncode == 0 and "elapsed time=" in r.stdout, r.stdout[-1500:]
    o = orc.Oracle(V, False, e, M, 1, 3, 1e-9, 0)
    o.initial_solve()
    assert np.abs(np.fromfile(dump, dtype=np.float64) - o.power_iteration()).max() <= 2e-9


def test_cli_multi_source(tmp_path):
    V, M = 3000, 30000
    e = graphgen.rmat_directed(V, M, seed=33)
    p = str(tmp_path / "g.bin"); dump = str(tmp_path / "p.f64"); sf = tmp_path / "src.txt"
    graphgen.write_bin(p, V, e)
    srcs = [int(x) for x in graphgen.top_out_degree(V, e, True, 4)]
    sf.write_text("\n".join(map(str, srcs)) + "\n")
    r = subprocess.run([CLI, "-d", p, "-i", "1", "-y", "1", "-n", "1", "-c", "50", "-l", "500", "--sources", str(sf),
                        "--dump", dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:]
    got = np.fromfile(dump, dtype=np.float64).reshape(4, V)
    wl = stream.workload(M, 0.1, 1, -1, 0, 50, 500)
    for i, s in enumerate(srcs):
        o = orc.Oracle(V, True, e, wl.W, wl.B, s, 1e-9, 0)
        o.initial_solve()
        for _ in range(wl.n_batches):
            o.slide(wl.B)
        assert np.abs(got[i] - o.p).max() <= 2e-9


def test_cli_validate_and_progress_cadence(tmp_path):
    """--validate 1 = the reference's -DVALIDATE checks on the device after every batch; the keys appear before EVERY batch
    when a batch holds more than 100 edges (gpu/PPRGPU.cuh:116), else before every 100th"""
    V, M = 6000, 50000
    e = graphgen.powerlaw_undirected(V, M, seed=31)
    p = str(tmp_path / "g.bin")
    graphgen.write_bin(p, V, e)
    base = [CLI, "-d", p, "-a", "0", "-i", "0", "-y", "1", "-w", "0.1", "-s", "7", "-e", "1e-9"]
    r = subprocess.run(base + ["-n", "0", "-r", "0.05", "-b", "6", "--validate", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    val = [ln for ln in r.stdout.splitlines() if ln.startswith("validate batch")]
    assert len(val) == 7 and all("window_mismatches=0 " in ln for ln in val)     # initial solve + 6 batches
    assert r.stdout.count("coming stream_batch_count=") == 7                       # B = 250 > 100: before each batch + the final block
    assert "VALIDATION FAILED" not in r.stdout
    r = subprocess.run(base + ["-n", "1", "-c", "10", "-l", "2500"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.count("coming stream_batch_count=") == 3                       # before batches 100, 200 + the final block
    r = subprocess.run(base + ["-n", "1", "-c", "10", "-l", "2500", "--progress", "0"], capture_output=True, text=True)
    assert r.stdout.count("coming stream_batch_count=") == 1
    assert int(_kv(r.stdout)["batches_done"]) == 250


def test_cli_pool_factor_flag_and_capacity_failure(tmp_path):
    """a pool that cannot hold a growing hub: the CLI fails loudly (round 1: the inserts were dropped silently)"""
    V = 2000
    rng = np.random.default_rng(1)
    win = rng.integers(0, V, size=(4000, 2)).astype(np.int32)
    hub = np.stack([rng.integers(0, V, size=3200), np.full(3200, 5)], axis=1).astype(np.int32)
    p = str(tmp_path / "g.bin")
    graphgen.write_bin(p, V, np.concatenate([win, hub] + [win] * 8)[:40000])   # window = first 4000 records, then 3200 edges into vertex 5
    args = [CLI, "-d", p, "-i", "1", "-y", "1", "-w", "0.1", "-n", "1", "-c", "400", "-l", "3200", "-s", "1"]
    ok = subprocess.run(args, capture_output=True, text=True)
    assert ok.returncode == 0, ok.stdout[-1500:]
    bad = subprocess.run(args + ["--pool-factor", "1.3"], capture_output=True, text=True)
    assert bad.returncode != 0 and "pool" in bad.stdout


def test_workload_tool_follows_the_reference_rule(tmp_path):
    """bin/workload against the reference's own tool (oracle/_ref/workload_ref): same file names; the degrees of the chosen
    ids agree position by position (same rule, same rand() draws; equal degrees may be different vertices: the reference's
    std::sort leaves ties in unspecified order)"""
    ref_tool = os.path.join(ROOT, "oracle", "_ref", "workload_ref")
    tool = os.path.join(ROOT, "dynamicppr_b200", "bin", "workload")
    V, M = 30_000, 400_000
    e = graphgen.powerlaw_undirected(V, M, seed=41)
    deg = np.bincount(e.ravel(), minlength=V)
    for sub in ("ours", "ref"):
        (tmp_path / sub).mkdir()
    p = str(tmp_path / "g.bin")
    graphgen.write_bin(p, V, e)
    r = subprocess.run([tool, p, "0", "0", "1"], capture_output=True, text=True, cwd=tmp_path / "ours")
    assert r.returncode == 0, r.stdout[-1500:]
    names = ["g.bin_top10.txt", "g.bin_top1000.txt", "g.bin_top1000000.txt"]
    order = np.lexsort((np.arange(V), -deg))
    rank = np.empty(V, np.int64); rank[order] = np.arange(V)
    ours = {}
    for nm, (lo, hi) in zip(names, [(0, 10), (10, 1000), (1000, min(1000000, V))]):
        ids = np.loadtxt(tmp_path / "ours" / nm, dtype=np.int64)
        assert len(ids) == 10 and len(set(ids.tolist())) == 10
        # every id sits in its bucket (by degree: ties may straddle the boundary only with equal degree)
        assert np.all(deg[ids] <= deg[order[lo]]) and np.all(deg[ids] >= deg[order[hi - 1]])
        ours[nm] = ids
    np.testing.assert_array_equal(ours[names[0]], order[:10])
    if os.path.exists(ref_tool):
        r = subprocess.run([ref_tool, p, "0", "0", "1"], capture_output=True, text=True, cwd=tmp_path / "ref")
        assert r.returncode == 0
        for nm in names:
            ref_ids = np.loadtxt(tmp_path / "ref" / nm, dtype=np.int64)
            np.testing.assert_array_equal(deg[ours[nm]], deg[ref_ids], err_msg=nm)


def test_cli_pick_chooses_the_fourth_id_of_the_bucket(tmp_path):
    V, M = 30_000, 400_000
    e = graphgen.powerlaw_undirected(V, M, seed=41)
    p = str(tmp_path / "g.bin")
    graphgen.write_bin(p, V, e)
    deg = np.bincount(e.ravel(), minlength=V)
    order = np.lexsort((np.arange(V), -deg))
    r = subprocess.run([CLI, "-d", p, "-i", "0", "-y", "1", "-w", "0.1", "-n", "0", "-r", "0.01", "-b", "2", "--pick", "top10"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-1500:]
    assert f"choose {order[3]} as source vertex id" in r.stdout
