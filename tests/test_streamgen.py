"""Seeded stream generators (include/dppr.h: dppr_generate_stream_host / _device, dppr_rank_by_degree) and the BASELINE
workload definitions built on them (dynamicppr_b200/workloads.py).

CPU part: the host twin is deterministic, slice-consistent (edge i depends on (kind, V, seed, i) only), produces valid
encoder-format payloads, and is pinned by digests committed here -- so the .bin prefix the reference CPU arm reads on
the GPU box is the stream the GPU arm generates.  GPU part: the device generator writes the same bytes as the host
twin; the device degree ranking equals numpy's."""
import hashlib
import os
import tempfile

import numpy as np
import pytest

from dynamicppr_b200 import binding, graphgen, workloads
from helpers import have_gpu

RMAT, PL = binding.STREAM_RMAT, binding.STREAM_POWERLAW


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


# first 4096 edges of the BASELINE streams: the generator must never drift (the committed profiles and the CPU arm depend on it)
PINNED = {
    (RMAT, 4_847_571, workloads.SEED + 3): "8d09baf2ff43b2c5",
    (PL, 3_072_441, workloads.SEED + 4): "23471935ad8091fe",
    (RMAT, 41_652_230, workloads.SEED + 5): "d88758314119c986",
}


@pytest.mark.parametrize("kind,V,seed", list(PINNED))
def test_host_generator_is_pinned(kind, V, seed):
    e = binding.generate_stream_host(kind, V, 0, 4096, seed, threads=3)
    assert digest(e) == PINNED[(kind, V, seed)]


@pytest.mark.parametrize("kind", [RMAT, PL])
def test_host_generator_slices_threads_and_ranges(kind):
    V, seed = 100_003, 77
    whole = binding.generate_stream_host(kind, V, 0, 300_000, seed, threads=1)
    assert whole.min() >= 0 and whole.max() < V
    for first, n, th in ((0, 300_000, 8), (12_345, 70_000, 3), (299_999, 1, 1), (5, 0, 2)):
        part = binding.generate_stream_host(kind, V, first, n, seed, threads=th)
        np.testing.assert_array_equal(part, whole[first:first + n])
    assert not np.array_equal(whole, binding.generate_stream_host(kind, V, 0, 300_000, seed + 1))
    if kind == PL:
        assert not np.any(whole[:, 0] == whole[:, 1]), "the power-law stream has no self-loops"
        deg = np.bincount(whole.ravel(), minlength=V)
        top = np.sort(deg)[::-1]
        assert top[0] > 50 * deg.mean() and top[0] < 0.02 * deg.sum()   # heavy tail, but no vertex owns the graph
        # rank k has weight ~ (k + 1)^-0.75: the ratio of the two heaviest vertices is 2^0.75 within sampling noise
        assert 1.3 < top[0] / top[1] < 2.2
    else:
        out = np.bincount(whole[:, 0], minlength=V)
        assert np.sort(out)[-1] > 100 * out.mean()


def test_bad_generator_arguments():
    with pytest.raises(binding.DpprError):
        binding.generate_stream_host(7, 100, 0, 10, 1)
    with pytest.raises(binding.DpprError):
        binding.generate_stream_host(RMAT, 1, 0, 10, 1)
    with pytest.raises(binding.DpprError):
        binding.generate_stream_host(RMAT, 100, -1, 10, 1)


def test_prefix_bin_is_sparse_full_size_and_reusable():
    cfg = workloads.scaled(workloads.CONFIGS[4], 0.002, n_sources=4)
    wl = cfg.workload()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "g.bin")
        n = wl.W + 3 * wl.B
        workloads.write_prefix_bin(cfg, path, n)
        assert os.path.getsize(path) == 4 + 8 * cfg.M              # the reference derives W from the file size
        V, edges = graphgen.read_bin(path)
        assert V == cfg.V and len(edges) == cfg.M
        np.testing.assert_array_equal(edges[:n], workloads.host_edges(cfg, 0, n))
        assert not edges[n:].any()                                  # the hole
        assert os.stat(path).st_blocks * 512 < os.path.getsize(path)
        mtime = os.stat(path).st_mtime_ns
        workloads.write_prefix_bin(cfg, path, n - 10)               # a shorter prefix is already there
        assert os.stat(path).st_mtime_ns == mtime
        # window arithmetic on the sparse file equals the config's
        from dynamicppr_b200 import stream
        assert stream.workload((os.path.getsize(path) - 4) // 8, 0.1, 0, 0.01, 100).W == wl.W


def test_baseline_configs_table():
    c = workloads.CONFIGS
    assert (c[4].workload().W, c[4].workload().B) == (11_718_508, 117_185)
    assert (c[5].workload().W, c[5].workload().B) == (146_836_518, 1_468_365)
    assert (c[3].workload().W, c[3].workload().B, c[3].workload().n_batches) == (6_899_377, 100, 100)
    assert c[3].cli_flags()[-4:] == ["-c", "100", "-l", "10000"]
    assert "-r" in c[4].cli_flags(7) and c[4].cli_flags(7)[-1] == "7"


def test_host_degree_order_matches_reference_rule():
    cfg = workloads.scaled(workloads.CONFIGS[4], 0.001)
    order = workloads.host_degree_order(cfg)
    e = workloads.host_edges(cfg, 0, cfg.M)
    deg = np.bincount(e.ravel(), minlength=cfg.V)               # undirected: an edge counts at both ends (workload/Graph.h:104-109)
    assert np.all(np.diff(deg[order]) <= 0)
    ties = np.diff(deg[order]) == 0
    assert np.all(np.diff(order)[ties] > 0)                      # ties by ascending id
    np.testing.assert_array_equal(workloads.top_sources(cfg, 10, on_host=True), order[:10])


@pytest.mark.gpu
@pytest.mark.parametrize("kind,V", [(RMAT, 4_847_571), (PL, 3_072_441), (PL, 1_000), (RMAT, 41_652_230)])
def test_device_generator_equals_host_twin(kind, V):
    import torch
    seed, first, n = 20261017 + kind, 123_456_789, 3_000_001
    t = torch.empty((n, 2), dtype=torch.int32, device="cuda")
    binding.generate_stream_device(kind, V, first, n, seed, t.data_ptr())
    np.testing.assert_array_equal(t.cpu().numpy(), binding.generate_stream_host(kind, V, first, n, seed))


@pytest.mark.gpu
def test_rmat_entry_point_is_the_rmat_stream():
    import torch
    t = torch.empty((10_000, 2), dtype=torch.int32, device="cuda")
    binding.generate_rmat_device(50_000, 10_000, 9, t.data_ptr())
    np.testing.assert_array_equal(t.cpu().numpy(), binding.generate_stream_host(RMAT, 50_000, 0, 10_000, 9))


@pytest.mark.gpu
@pytest.mark.parametrize("directed", [True, False])
def test_device_degree_ranking_is_exact(directed):
    import torch
    cfg = workloads.scaled(workloads.CONFIGS[3 if directed else 4], 0.01)
    e = workloads.host_edges(cfg, 0, cfg.M)
    out = np.bincount(e[:, 0], minlength=cfg.V); inn = np.bincount(e[:, 1], minlength=cfg.V)
    if not directed:
        out = out + inn; inn = out
    want = np.lexsort((np.arange(cfg.V), -out)).astype(np.int32)
    order, od, idg = binding.rank_by_degree(cfg.V, directed, pairs=e, want_degrees=True)          # host pairs, chunked upload
    np.testing.assert_array_equal(order, want)
    np.testing.assert_array_equal(od, out); np.testing.assert_array_equal(idg, inn)
    t = torch.from_numpy(e).cuda()
    np.testing.assert_array_equal(binding.rank_by_degree(cfg.V, directed, n_edges=len(e), device_ptr=t.data_ptr()), want)
    by_in = binding.rank_by_degree(cfg.V, directed, pairs=e, by_out_degree=False)
    np.testing.assert_array_equal(by_in, np.lexsort((np.arange(cfg.V), -inn)).astype(np.int32))
    np.testing.assert_array_equal(workloads.top_sources(cfg, 25), want[:25])                      # device path of the workloads
    with pytest.raises(binding.DpprError):
        bad = e[:100].copy(); bad[5, 1] = cfg.V
        binding.rank_by_degree(cfg.V, directed, pairs=bad)
