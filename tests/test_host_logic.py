"""CPU tests: host logic, the C-ABI library's exported surface, the CLI contract, the .bin format,
the multi-rank plumbing (gloo, world_size 2).  No compute call needs a GPU here."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import oracle as orc
from dynamicppr_b200 import binding, graphgen, sharding, stream
from helpers import GOLDEN, GOLDEN_IDS, golden_workload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "dynamicppr_b200", "bin", "pagerank")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not (os.path.exists(binding.library_path()) and os.path.exists(CLI)):
        subprocess.check_call(["make", "-s", "lib", "cli"], cwd=ROOT)


def test_library_exports_every_symbol_the_header_declares():
    hdr = open(os.path.join(ROOT, "include", "dppr.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(dppr_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed from include/dppr.h"
    assert declared == set(binding.ABI_SYMBOLS), (declared ^ set(binding.ABI_SYMBOLS))
    lib = binding.load_library()
    for name in sorted(declared):
        assert getattr(lib, name) is not None
    assert lib.dppr_version() == 200


def test_create_fails_loudly_without_a_gpu_or_with_bad_config():
    from helpers import have_gpu
    from dynamicppr_b200 import DynamicPPR, DpprError
    with pytest.raises(DpprError) as ei:
        DynamicPPR(0, True, 10, 1, [0])
    assert ei.value.code == 1 and "vertex_count" in str(ei.value)
    if not have_gpu():
        with pytest.raises(DpprError) as ei:
            DynamicPPR(10, True, 5, 1, [0])
        assert ei.value.code == 5 and "no CPU fallback" in str(ei.value)


def test_struct_layouts_match_the_header():
    # sizes the C side computes for the same declarations (compiled on the fly with gcc)
    src = '#include <stdio.h>\n#include "dppr.h"\nint main(){printf("%zu %zu\\n", sizeof(dppr_config), sizeof(dppr_batch_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", os.path.join(d, "t")])
        a, b = map(int, subprocess.check_output([os.path.join(d, "t")]).split())
    assert a == ctypes.sizeof(binding.Config) and b == ctypes.sizeof(binding.BatchStats)


@pytest.mark.parametrize("path", GOLDEN, ids=GOLDEN_IDS)
def test_python_stream_arithmetic_matches_reference_dump(path):
    g = np.load(path)
    wl = golden_workload(g)
    assert wl.W == int(g["W"]) and wl.B == int(g["B"])
    assert wl.runnable_batches(len(g["edges"])) + 1 == int(g["n_snap"])


@pytest.mark.parametrize("M,wr,mode,r,b,c,l", [(1_049_866, 0.1, 0, 0.01, 100, 0, 0), (2_987_624, 0.1, 0, 0.01, 100, 0, 0),
                                                (68_993_773, 0.1, 1, -1, 0, 100, 10_000), (117_185_083, 0.1, 0, 0.01, 100, 0, 0),
                                                (1_468_365_182, 0.1, 0, 0.01, 100, 0, 0), (1000, 0.5, 0, 0.3, 10, 0, 0),
                                                (999, 0.37, 1, -1, 0, 7, 1000)])
def test_stream_arithmetic_matches_oracle_on_baseline_shapes(M, wr, mode, r, b, c, l):
    W, B, nb, tot = orc.workload(M, wr, mode, r, b, c, l)
    wl = stream.workload(M, wr, mode, r, b, c, l)
    assert (wl.W, wl.B, wl.n_batches, wl.total) == (W, B, nb, tot)


def test_baseline_table_values():
    """SURVEY section 8 table"""
    assert stream.workload(1_049_866, 0.1, 0, 0.01, 100).B == 1049
    wl = stream.workload(2_987_624, 0.1, 0, 0.01, 100)
    assert (wl.W, wl.B) == (298_762, 2_987)
    wl = stream.workload(1_468_365_182, 0.1, 0, 0.01, 100)
    assert (wl.W, wl.B) == (146_836_518, 1_468_365)


def test_bin_round_trip_and_generators(tmp_path):
    e = graphgen.powerlaw_undirected(500, 3000, seed=5)
    assert e.shape == (3000, 2) and e.dtype == np.int32 and e.min() >= 0 and e.max() < 500
    assert (e[:, 0] != e[:, 1]).all()
    key = np.minimum(e[:, 0], e[:, 1]).astype(np.int64) * 500 + np.maximum(e[:, 0], e[:, 1])
    assert len(np.unique(key)) == 3000  # no duplicate undirected pair
    np.testing.assert_array_equal(e, graphgen.powerlaw_undirected(500, 3000, seed=5))  # seeded
    d = graphgen.rmat_directed(1000, 5000, seed=6)
    assert d.shape == (5000, 2) and d.min() >= 0 and d.max() < 1000
    p = str(tmp_path / "g.bin")
    graphgen.write_bin(p, 500, e)
    assert os.path.getsize(p) == 4 + 8 * 3000
    V, back = graphgen.read_bin(p)
    assert V == 500
    np.testing.assert_array_equal(np.asarray(back), e)


def test_reference_encoder_output_is_readable(tmp_path):
    """the .bin the reference's own encoder writes (encoder/GraphEncoder.h:86-95) parses with our reader"""
    enc = os.path.join(ROOT, "oracle", "_ref", "encoder_ref")
    if not os.path.exists(enc):
        pytest.skip("oracle/_ref/encoder_ref not built (needs /root/reference)")
    rng = np.random.default_rng(0)
    e = rng.integers(3, 50, size=(200, 2))
    txt = tmp_path / "toy.txt"
    txt.write_text("".join(f"{a}\t{b}\n" for a, b in e))
    subprocess.check_call([enc, str(txt)], cwd=tmp_path, stdout=subprocess.DEVNULL)
    V, back = graphgen.read_bin(str(tmp_path / "toy.bin"))
    assert V == e.max() - e.min() + 1
    got = np.asarray(back)
    assert got.shape == (200, 2)
    # the encoder shuffles record order (rand()) and rebases ids to 0: compare as multisets
    exp = (e - e.min()).astype(np.int32)
    assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, exp.tolist()))


def test_top_out_degree_matches_reference_workload_tool(tmp_path):
    tool = os.path.join(ROOT, "oracle", "_ref", "workload_ref")
    if not os.path.exists(tool):
        pytest.skip("oracle/_ref/workload_ref not built (needs /root/reference)")
    e = graphgen.powerlaw_undirected(400, 4000, seed=9)
    p = str(tmp_path / "w.bin")
    graphgen.write_bin(p, 400, e)
    subprocess.check_call([tool, p, "0", "0", "1"], cwd=tmp_path, stdout=subprocess.DEVNULL)
    ref = [int(x) for x in open(tmp_path / "w.bin_top10.txt").read().split()]
    deg = np.bincount(e[:, 0], minlength=400) + np.bincount(e[:, 1], minlength=400)
    ours = graphgen.top_out_degree(400, e, False, 10)
    # same degree sequence (the reference's std::sort is unstable, so ties may be ordered differently)
    assert sorted(deg[ref].tolist(), reverse=True) == deg[ours].tolist()


# ---- CLI contract (Arguments.h:42-86): validation and echo work without a GPU ----------------------
def _cli(*args):
    return subprocess.run([CLI, *args], capture_output=True, text=True)


def test_cli_rejects_invalid_arguments_like_the_reference():
    for args in ([], ["-d", "x.bin"], ["-d", "x.bin", "-i", "0", "-y", "1"],            # mode 0 needs -r and -b
                 ["-d", "x.bin", "-i", "0", "-y", "1", "-n", "1", "-c", "10"],           # mode 1 needs -l
                 ["-d", "x.bin", "-i", "0", "-y", "1", "-n", "2", "-r", "0.1", "-b", "3"],
                 ["-d", "x.bin", "-i", "0", "-y", "1", "-r", "0.1", "-b", "3", "-o", "9"]):
        r = _cli(*args)
        assert r.returncode == 255, args  # exit(-1)
        assert r.stdout.startswith("invalid arguments\n==========[USAGE]==========")


def test_cli_echoes_arguments_and_workload_lines_like_the_reference(tmp_path):
    e = graphgen.powerlaw_undirected(300, 3000, seed=11)
    p = str(tmp_path / "g.bin")
    graphgen.write_bin(p, 300, e)
    r = _cli("-d", p, "-a", "0", "-i", "0", "-y", "1", "-w", "0.1", "-n", "0", "-r", "0.07", "-b", "25", "-s", "7")
    out = r.stdout
    assert "gAppType=0,gIsDirected=0,gIsDynamic=1\n" in out
    assert "gWindowRatio=0.1,gWorkloadConfigType=0,gStreamUpdateCountVersusWindowRatio=0.07,gStreamBatchCount=25," in out
    assert "gSourceVertexId=7\n" in out and "error=1e-09,ALPHA=0.15\n" in out
    assert "vertex_count=300\n" in out
    assert "after workload config: gStreamUpdateCountPerBatch=21,gStreamBatchCount=25,gStreamUpdateCountTotal=525\n" in out
    assert "sliding window size=300,gStreamUpdateCountPerBatch=21\n" in out and "edge_count=600\n" in out
    from helpers import have_gpu
    if not have_gpu():
        assert r.returncode == 255 and "no CPU fallback" in out  # fails loudly, exit(-1) like CUDA_ERROR


def test_extract_gpu_rule_parses_our_keys():
    """scripts/extract_gpu.py:18-32: last line containing the keyword, token after the first space"""
    sample = ("batches_done 3\np50_batch_ms 1.5\ncoming stream_batch_count=4\nppr_time 4.5\nedge_num 300\n"
              "ppr_latency 1.5\nppr_throughput 66666.7\n")

    def extract(text, keyword):
        val = ""
        for line in text.splitlines(True):
            pos = line.find(keyword)
            if pos >= 0:
                s = line.find(" ", pos); t = line.find(" ", s + 1)
                tok = line[s + 1: t]
                try:
                    float(tok); val = tok
                except ValueError:
                    pass
        return float(val)

    assert extract(sample, "ppr_latency") == 1.5 and extract(sample, "ppr_throughput") == 66666.7
    src = open(os.path.join(ROOT, "dynamicppr_b200", "host", "PPRDriver.h")).read()
    extra_keys = re.findall(r'std::cout << "([a-z0-9_]+) "', src)
    assert "ppr_latency" in extra_keys and "ppr_throughput" in extra_keys
    for k in extra_keys:  # no other key may contain the scraped substrings
        if k not in ("ppr_latency", "ppr_throughput"):
            assert "ppr_latency" not in k and "ppr_throughput" not in k


# ---- N > 1 plumbing on CPU (gloo, world_size 2) ---------------------------------------------------
def test_shard_sources():
    s = np.arange(10, dtype=np.int32)
    assert sharding.shard_sources(s, 1, 4, 2).tolist() == [2, 3]
    parts = [sharding.shard_sources(s, r, 4).tolist() for r in range(4)]
    assert sum(parts, []) == list(range(10)) and [len(p) for p in parts] == [3, 3, 2, 2]
    with pytest.raises(ValueError):
        sharding.shard_sources(s, 0, 8, 2)
    dealt = [sharding.shard_sources(s, r, 4, interleave=True).tolist() for r in range(4)]
    assert dealt == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    with pytest.raises(ValueError):
        sharding.shard_sources(s, 0, 2, 2, interleave=True)


_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], 'tests'))
from dynamicppr_b200 import sharding, graphgen, stream
import oracle as orc
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dist.init_process_group('gloo', rank=rank, world_size=world)
V, M = 400, 4000
edges = graphgen.powerlaw_undirected(V, M, seed=3)           # every rank replicates the stream
wl = stream.workload(M, 0.1, 0, 0.05, 4)
sources = graphgen.top_out_degree(V, edges, False, 3)       # 3 sources over 2 ranks: ragged shards
mine = sharding.shard_sources(sources, rank, world)
rows = []
for s in mine:                                               # the oracle stands in for the engine on CPU
    o = orc.Oracle(V, False, edges, wl.W, wl.B, int(s), 1e-9, 0); o.initial_solve()
    for _ in range(wl.n_batches): o.slide(wl.B)
    rows.append(o.p)
local = torch.from_numpy(np.stack(rows))
t = sharding.max_over_ranks([1.0 + rank, 5.0 - rank])
assert t == [float(world), 5.0], t
out = sharding.gather_estimates(local, dst=0)
if rank == 0:
    allp = torch.cat(out).numpy()
    assert allp.shape == (3, V)
    for i, s in enumerate(sources):                          # equals a single-process run of every source
        o = orc.Oracle(V, False, edges, wl.W, wl.B, int(s), 1e-9, 0); o.initial_solve()
        for _ in range(wl.n_batches): o.slide(wl.B)
        assert np.array_equal(allp[i], o.p), i
    print('GATHER_OK')
dist.destroy_process_group()
"""


def test_two_rank_gloo_source_sharding_and_gather(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(w), ROOT],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "GATHER_OK" in r.stdout


def test_internal_vertex_order_is_a_bijection_with_hot_block_prefixes():
    """csrc/window.cuh relabel_slot: rank k (0 = highest out-degree) -> internal id.  Host arithmetic, no GPU."""
    L = binding.load_library()
    for V, P in [(1, 1), (7, 3), (10, 1024), (1000, 1), (1000, 7), (4097, 1024), (100_003, 1024)]:
        P = min(P, V)
        ids = np.array([L.dppr_test_relabel_slot(k, V, P) for k in range(V)], dtype=np.int64)
        assert sorted(ids.tolist()) == list(range(V)), (V, P)           # a permutation of [0, V)
        q, rem = divmod(V, P)
        starts = np.array([r * q + min(r, rem) for r in range(P)])      # first id of each block
        np.testing.assert_array_equal(ids[:P], starts)                  # the P hottest vertices head the P blocks
        if V >= 2 * P:                                                   # ... and consecutive ranks are a block apart
            assert np.all(np.abs(np.diff(ids[:P])) >= q)
            np.testing.assert_array_equal(ids[P: 2 * P], starts + 1)    # the next P sit right behind them
