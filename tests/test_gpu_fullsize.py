"""GPU: BASELINE.json configs at FULL size, checked through size-independent properties (no oracle run needed):
  * canonical window CSR bit-exact against a numpy lexsort of the same window of the stream,
  * residual bound |r| <= eps,
  * the push invariant  p[u] + a r[u] = a [u==s] + (1-a)/(outdeg(u)+1) * sum_{w in out(u)} p[w]  for EVERY vertex
     -- together with the residual bound it implies |p - pi| <= eps, i.e. the 2-eps parity criterion.
Config 1 (dblp-shaped) and config 2 (youtube-shaped) are checked against numpy on the host; configs 3 (all four
variants), 4 (8 of the top-1000 sources) and 5 (one source, a few batches) run at FULL size with every check on the
device: window graph bit-exact against the canonical entry list built from the window's own edges
(dppr_check_window_device), residual bound and push invariant by dppr_validate -- SURVEY Appendix E T3-T6, T10."""
import numpy as np
import pytest

from dynamicppr_b200 import DynamicPPR, graphgen, stream, workloads

pytestmark = pytest.mark.gpu
ALPHA = 0.15


def numpy_window_csr(V, directed, win):
    dst, src = win[:, 1].astype(np.int64), win[:, 0].astype(np.int64)
    if not directed:
        dst, src = np.concatenate([dst, win[:, 0]]), np.concatenate([src, win[:, 1]])
    order = np.lexsort((src, dst))
    rp = np.zeros(V + 1, np.int64)
    np.cumsum(np.bincount(dst, minlength=V), out=rp[1:])
    return rp.astype(np.int32), src[order].astype(np.int32), np.bincount(src, minlength=V).astype(np.int32)


def invariant_defect(V, rp, ci, od, p, r, source):
    """max_u | p + a r - a e_s - (1-a)/(d+1) * sum_{w in out(u)} p[w] |  from the exported in-CSR"""
    indeg = np.diff(rp)
    acc = np.bincount(ci, weights=np.repeat(p, indeg), minlength=V)  # edge u->w sits in row w, column u
    lhs = p + ALPHA * r
    lhs[source] -= ALPHA
    return np.abs(lhs - (1 - ALPHA) * acc / (od + 1.0)).max()


@pytest.mark.parametrize("shape,source_kind,batches", [("dblp", "s1", 100), ("youtube", "top", 100)])
def test_baseline_config_full_size(shape, source_kind, batches):
    V, M, directed = graphgen.SHAPES[shape]
    edges = graphgen.powerlaw_undirected(V, M, graphgen.BASE_SEED + list(graphgen.SHAPES).index(shape))
    wl = stream.workload(M, 0.1, 0, 0.01, batches)
    src = 1 if source_kind == "s1" else int(graphgen.top_out_degree(V, edges, directed, 1)[0])
    eps = 1e-9
    with DynamicPPR(V, directed, wl.W, wl.B, [src], epsilon=eps) as eng:
        eng.init_window_pairs(edges[: wl.W])
        eng.solve_initial()
        for k in range(batches + 1):
            if k > 0:
                lo = wl.W + (k - 1) * wl.B
                eng.slide_pairs(edges[lo: lo + wl.B])
            if k not in (0, 1, batches // 2, batches):
                continue
            assert eng.stats().error_flags == 0
            rp, ci, od = eng.export_window_csr()
            erp, eci, eod = numpy_window_csr(V, directed, edges[k * wl.B: k * wl.B + wl.W])
            np.testing.assert_array_equal(rp, erp)
            np.testing.assert_array_equal(ci, eci)
            np.testing.assert_array_equal(od, eod)
            p, r = eng.estimates(), eng.residuals()
            assert np.abs(r).max() <= eps
            assert invariant_defect(V, rp, ci, od, p, r, src) <= 1e-13
            assert abs(p.sum()) < 1e6 and np.all(np.isfinite(p))


@pytest.mark.parametrize("directed", [False, True], ids=["orkut-shaped", "livejournal-shaped"])
def test_multi_source_default_switching_kernel(directed):
    """8 top-degree sources on a window large enough (>= 2e7 directed edges x sources) that the engine picks, by itself
    and with default settings, the kernel that switches between scatter iterations and gather sweeps (csrc/pull.cuh):
    BASELINE configs[3] / configs[2] shapes at 1/8 and 1/2 scale.  Same size-independent checks, for every source."""
    if directed:
        V, M = 2_423_785, 34_496_886
        edges = graphgen.rmat_directed(V, M, graphgen.BASE_SEED + 2)
    else:
        V, M = 384_055, 14_648_135
        edges = graphgen.powerlaw_undirected(V, M, graphgen.BASE_SEED + 3)
    batches = 3
    wl = stream.workload(M, 0.1, 0, 0.01, batches)
    assert wl.W * (1 if directed else 2) * 8 >= 2e7
    sources = [int(x) for x in graphgen.top_out_degree(V, edges, directed, 8)]
    eps = 1e-9
    sweeps = 0
    with DynamicPPR(V, directed, wl.W, wl.B, sources, epsilon=eps) as eng:
        eng.init_window_pairs(edges[: wl.W])
        eng.solve_initial()
        for k in range(batches + 1):
            if k > 0:
                lo = wl.W + (k - 1) * wl.B
                eng.slide_pairs(edges[lo: lo + wl.B])
            st = eng.stats()
            assert st.error_flags == 0
            sweeps += st.dense_sweeps
            if k not in (0, batches):
                continue
            rp, ci, od = eng.export_window_csr()
            erp, eci, eod = numpy_window_csr(V, directed, edges[k * wl.B: k * wl.B + wl.W])
            np.testing.assert_array_equal(rp, erp)
            np.testing.assert_array_equal(ci, eci)
            np.testing.assert_array_equal(od, eod)
            if directed:  # the out-lists the sweeps read are the transpose of the exported in-lists
                out = eng.export_window_out_csr()
                assert out is not None
                order = np.lexsort((np.repeat(np.arange(V, dtype=np.int64), np.diff(rp)), ci.astype(np.int64)))
                np.testing.assert_array_equal(out[1], np.repeat(np.arange(V, dtype=np.int32), np.diff(rp))[order])
                np.testing.assert_array_equal(np.diff(out[0]), od)
            for i, s in enumerate(sources):
                p, r = eng.estimates(i), eng.residuals(i)
                assert np.abs(r).max() <= eps, (k, i)
                assert invariant_defect(V, rp, ci, od, p, r, s) <= 1e-13, (k, i)
    if not directed:  # (on the directed shape the device-side cost model may legitimately keep scattering)
        assert sweeps > 0, "the gather sweeps never ran: the default cost model or size threshold changed"


def _fullsize_run(cfg, sources, batches, variant=0, check_at=None, **kw):
    """stream generated on the device, engine fed device pointers, every check on the device"""
    import torch
    wl = cfg.workload()
    batches = min(batches, wl.runnable_batches(cfg.M))
    check_at = set(check_at if check_at is not None else (0, batches))
    dev = workloads.device_edges(cfg, 0, wl.W + batches * wl.B)
    stats = []
    with DynamicPPR(cfg.V, cfg.directed, wl.W, wl.B, sources, epsilon=cfg.eps, variant=variant, **kw) as eng:
        eng.init_window_device_pairs(dev.data_ptr(), wl.W)
        eng.solve_initial()
        for k in range(batches + 1):
            if k > 0:
                eng.slide_device_pairs(dev.data_ptr() + 8 * (wl.W + (k - 1) * wl.B), wl.B)
            if k not in check_at:
                continue
            st = eng.stats()
            assert st.error_flags == 0 and st.pool_leaked == 0, (k, st.error_flags)
            stats.append(st)
            assert eng.check_window_device(dev.data_ptr() + 8 * k * wl.B, wl.W) == 0, f"window graph differs after batch {k}"
            for i in range(len(sources)):
                max_r, defect = eng.validate(i)
                assert max_r <= cfg.eps, (k, i, max_r)
                assert defect <= 1e-13, (k, i, defect)
        ids, vals = eng.topk(4)
        for i, s in enumerate(sources):
            assert int(s) in ids[i], "a source is among the top estimates of its own PPR vector"
    del dev
    torch.cuda.empty_cache()
    return stats


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_config3_livejournal_shaped_mode1_all_variants_full_size(variant):
    """BASELINE configs[2]: 4.8 M vertices, 69 M-edge R-MAT stream, -n 1 -c 100 -l 10000 (100 batches of 100 edges)"""
    cfg = workloads.CONFIGS[3]
    src = workloads.top_sources(cfg, 1)
    stats = _fullsize_run(cfg, src, 100, variant=variant, check_at=(0, 1, 50, 100))
    assert stats[-1].batch_index == 100 and stats[-1].edges == 100


def test_config4_orkut_shaped_multi_source_full_size():
    """BASELINE configs[3]: 3.07 M vertices, 117 M-edge undirected stream; 8 of the job's top-1000 sources (list ranks
    0, 142, ..., 999) in one engine -- the switching kernel with gather sweeps"""
    cfg = workloads.CONFIGS[4]
    job = workloads.top_sources(cfg, 1000)
    srcs = job[np.linspace(0, 999, 8).astype(int)]
    stats = _fullsize_run(cfg, srcs, 4, check_at=(0, 2, 4))
    assert sum(st.dense_sweeps for st in stats) > 0, "the gather sweeps never ran"
    assert stats[-1].dense_pairs > 0 and stats[-1].scatter_edges > 0


def test_config4_orkut_shaped_source_panels_full_size():
    """the same graph, 40 sources refreshed in 3 panels (tuning.panel_sources = 16 -> 14 + 13 + 13; the default of 128 is what
    bench.py's 1000 sources run with): every panel is its own [V][16] block and its own repair / push launches, and every
    source of every panel satisfies the residual bound and the push invariant"""
    cfg = workloads.CONFIGS[4]
    job = workloads.top_sources(cfg, 1000)
    srcs = job[np.linspace(0, 999, 40).astype(int)]
    stats = _fullsize_run(cfg, srcs, 3, check_at=(0, 3), tuning={"panel_sources": 16})
    assert stats[-1].dense_sweeps > 3 * 10, "three panels, each with its own sweep episode"


def test_config5_twitter_shaped_one_source_full_size():
    """BASELINE configs[4] shape on one GPU: 41.7 M vertices, 146.8 M-edge window, 1.47 M-edge batches; the top source;
    window graph compared bit-exactly on the device (round 1 used a checksum)"""
    cfg = workloads.CONFIGS[5]
    src = workloads.top_sources(cfg, 1)
    stats = _fullsize_run(cfg, src, 3, check_at=(0, 3))
    assert stats[-1].edges == 1_468_365 and stats[-1].traversed_edges > 1e9
