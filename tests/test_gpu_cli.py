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
