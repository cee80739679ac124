"""GPU parity tests of the DENSE iterations (csrc/pull.cuh): when the frontier is large, variant 0 of the
level-synchronous engine runs an iteration as a gather sweep over the out-lists instead of scattering atomics.
Same contract as everywhere else: window graph bit-exact, estimates within 2 eps of the reference CPU push and of
power iteration, residuals within eps.  The thresholds are forced down through the environment so that tiny graphs
exercise every tier (lane group / warp / grid) and the switch in both directions."""
import numpy as np
import pytest

import oracle as orc
from dynamicppr_b200 import DynamicPPR, graphgen, stream, binding
from helpers import GOLDEN, GOLDEN_IDS, golden_workload, d2_possible, check_against

pytestmark = pytest.mark.gpu


def transpose_csr(V, rp, ci):
    """in-CSR (rows = destination, sorted by (dst, src), duplicates kept) -> out-CSR of the same multigraph"""
    dst = np.repeat(np.arange(V, dtype=np.int64), np.diff(rp.astype(np.int64)))
    src = ci.astype(np.int64)
    order = np.lexsort((dst, src))
    out_rp = np.zeros(V + 1, np.int64)
    np.cumsum(np.bincount(src, minlength=V), out=out_rp[1:])
    return out_rp.astype(np.int32), dst[order].astype(np.int32)


def force_dense(monkeypatch, div="1e15", tiers=(3, 12, 5)):
    """tiers = (warp_min, big_min, big_chunk): out-lists from warp_min entries are walked by a whole warp, from big_min
    entries they are cut into chunks of big_chunk entries that any warp of the grid takes (csrc/pull.cuh)"""
    monkeypatch.setenv("DPPR_DENSE_DIV", div)        # 1e15: enter at any frontier size >= 1
    monkeypatch.setenv("DPPR_DENSE_MIN_EDGES", "0")  # (by default small windows never switch)
    for name, v in zip(("WARP_MIN", "BIG_MIN", "BIG_CHUNK"), tiers):
        monkeypatch.setenv(f"DPPR_PULL_{name}", str(v))


def check_out_lists(eng, V, rp, ci, directed, tag):
    out = eng.export_window_out_csr()
    if not directed:
        assert out is None, tag  # undirected: the in-lists are the out-lists
        return
    assert out is not None, tag
    erp, eci = transpose_csr(V, rp, ci)
    np.testing.assert_array_equal(out[0], erp, err_msg=tag + " (out row_ptr)")
    np.testing.assert_array_equal(out[1], eci, err_msg=tag + " (out col_ind)")


@pytest.mark.parametrize("tiers", [(3, 12, 5), (2, 2, 1), (1000000, 1000000, 1000000), (1, 1000000, 1000000), (4, 4, 3)],
                         ids=["all-tiers", "grid-tier-chunks-of-1", "lane-tier", "warp-tier", "grid-tier"])
@pytest.mark.parametrize("path", GOLDEN, ids=GOLDEN_IDS)
def test_golden_with_forced_dense_iterations(path, tiers, monkeypatch):
    _golden_forced_dense(path, tiers, 0, monkeypatch)


@pytest.mark.parametrize("div", ["1e15", "64"], ids=["always", "mixed"])
@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("path", [p for p in GOLDEN if any(t in p for t in ("pl_undirected", "hub_expiry", "dense_multi_directed", "rmat_directed"))],
                         ids=lambda p: p.split("/")[-1][:-4])
def test_golden_dense_iterations_for_the_other_variants(path, variant, div, monkeypatch):
    """the sweeps are the same for every variant (they differ in how the SCATTER form reads residuals and dedupes); scatter
    iterations of variants 1-3 (snapshot passes, status stamps, repair pass) alternate with gather episodes"""
    _golden_forced_dense(path, (3, 12, 5), variant, monkeypatch, div=div)


def _golden_forced_dense(path, tiers, variant, monkeypatch, div="1e15"):
    force_dense(monkeypatch, div=div, tiers=tiers)
    g = np.load(path)
    V, directed, edges = int(g["V"]), bool(g["directed"]), g["edges"]
    wl = golden_workload(g)
    eps = float(g["eps"])
    use_ref_p = not d2_possible(g)
    sweeps = 0
    with DynamicPPR(V, directed, wl.W, wl.B, [int(g["source"])], epsilon=eps, variant=variant,
                    engine_mode=binding.ENGINE_LEVELSYNC) as eng:
        eng.init_window_pairs(edges[: wl.W])
        eng.solve_initial()
        for k in range(int(g["n_snap"])):
            if k > 0:
                lo = wl.W + (k - 1) * wl.B
                eng.slide_pairs(edges[lo: lo + wl.B])
            tag = f"{path} dense tiers {tiers} snapshot {k}"
            st = eng.stats()
            assert st.error_flags == 0, tag
            sweeps += st.dense_sweeps
            assert st.dense_sweeps <= st.iterations, tag
            rp, ci, od = eng.export_window_csr()
            np.testing.assert_array_equal(rp, g["in_row_ptr"][k], err_msg=tag)
            np.testing.assert_array_equal(ci, g["in_col"][k], err_msg=tag)
            np.testing.assert_array_equal(od, g["outdeg"][k], err_msg=tag)
            check_out_lists(eng, V, rp, ci, directed, tag)
            ref_p = g[f"v{variant}_p"][k] if (use_ref_p and f"v{variant}_p" in g) else None
            check_against(eng.estimates(), eng.residuals(), ref_p, g["pow"][k], eps, tag)
    assert sweeps > 0, f"{path}: the dense path never ran"


def _oracle_vs_dense(V, directed, edges, wl, source, eps, n_batches, check_every=1):
    o = orc.Oracle(V, directed, edges, wl.W, wl.B, source, eps, 0)
    o.initial_solve()
    sweeps = iters = 0
    with DynamicPPR(V, directed, wl.W, wl.B, [source], epsilon=eps, variant=0, engine_mode=binding.ENGINE_LEVELSYNC) as eng:
        eng.init_window_pairs(edges[: wl.W])
        eng.solve_initial()
        for k in range(n_batches + 1):
            if k > 0:
                lo = wl.W + (k - 1) * wl.B
                assert o.slide(wl.B) == 0
                eng.slide_pairs(edges[lo: lo + wl.B])
            st = eng.stats()
            tag = f"V={V} dense batch {k}"
            assert st.error_flags == 0, tag
            sweeps += st.dense_sweeps
            iters += st.iterations
            if k % check_every and k != n_batches:
                continue
            rp, ci, od = eng.export_window_csr()
            erp, eci, eod = o.canonical_csr()
            np.testing.assert_array_equal(rp, erp, err_msg=tag)
            np.testing.assert_array_equal(ci, eci, err_msg=tag)
            np.testing.assert_array_equal(od, eod, err_msg=tag)
            check_out_lists(eng, V, rp, ci, directed, tag)
            pw = o.power_iteration() if (k in (0, n_batches)) else None
            check_against(eng.estimates(), eng.residuals(), o.p, pw, eps, tag)
    return sweeps, iters


@pytest.mark.parametrize("div", ["1e15", "8", "64"], ids=["always", "default", "early"])
def test_top_degree_source_undirected_dense(div, monkeypatch):
    """the heavy-push case of test_gpu_parity (tens of iterations per batch), sweeps and scatter iterations mixed"""
    monkeypatch.setenv("DPPR_DENSE_DIV", div)
    monkeypatch.setenv("DPPR_DENSE_MIN_EDGES", "0")
    V, M, directed = 39_635, 131_233, False
    edges = graphgen.powerlaw_undirected(V, M, seed=graphgen.BASE_SEED)
    src = int(graphgen.top_out_degree(V, edges, directed, 1)[0])
    wl = stream.workload(M, 0.1, 0, 0.01, 100)
    sweeps, iters = _oracle_vs_dense(V, directed, edges, wl, src, 1e-9, 10, check_every=5)
    assert iters > 50
    if div != "8":
        assert sweeps > 0
    if div == "64":
        assert sweeps < iters  # both kinds of iteration ran: the switch was taken in both directions


@pytest.mark.parametrize("window_path", ["fused", "coop", "multikernel"])
def test_directed_out_lists_follow_the_window(window_path, monkeypatch):
    """directed R-MAT stream: the out-lists are maintained by each of the three launch shapes of the window update;
    the dense sweeps read them"""
    force_dense(monkeypatch, div="256", tiers=(8, 512, 128))
    if window_path != "fused":
        monkeypatch.setenv("DPPR_FUSED_WINDOW", "0")
    if window_path == "multikernel":
        monkeypatch.setenv("DPPR_COOP_WINDOW", "0")
    V, M, directed = 20_000, 400_000, True
    edges = graphgen.rmat_directed(V, M, seed=graphgen.BASE_SEED + 5)
    src = int(graphgen.top_out_degree(V, edges, directed, 1)[0])
    wl = stream.workload(M, 0.1, 1, -1.0, 0, 300, 300 * 12)
    sweeps, _ = _oracle_vs_dense(V, directed, edges, wl, src, 1e-9, 12, check_every=4)
    assert sweeps > 0


def test_multi_source_dense_matches_single_source_oracles(monkeypatch):
    """7 sources -> rows of 8 in x (one padding column), two source chunks per vertex"""
    force_dense(monkeypatch, div="1e15", tiers=(4, 200, 64))
    for directed in (False, True):
        V, M = 6_000, 50_000
        edges = graphgen.rmat_directed(V, M, seed=5) if directed else graphgen.powerlaw_undirected(V, M, seed=77)
        wl = stream.workload(M, 0.1, 0, 0.02, 6)
        sources = [int(x) for x in graphgen.top_out_degree(V, edges, directed, 5)] + [1, 2]
        eps = 1e-9
        oracles = []
        for s in sources:
            o = orc.Oracle(V, directed, edges, wl.W, wl.B, s, eps, 0)
            o.initial_solve()
            oracles.append(o)
        sweeps = 0
        with DynamicPPR(V, directed, wl.W, wl.B, sources, epsilon=eps) as eng:
            eng.init_window_pairs(edges[: wl.W])
            eng.solve_initial()
            for k in range(wl.n_batches + 1):
                if k > 0:
                    lo = wl.W + (k - 1) * wl.B
                    eng.slide_pairs(edges[lo: lo + wl.B])
                    for o in oracles:
                        o.slide(wl.B)
                st = eng.stats()
                assert st.error_flags == 0
                sweeps += st.dense_sweeps
                for i, o in enumerate(oracles):
                    check_against(eng.estimates(i), eng.residuals(i), o.p, None, eps,
                                  f"directed={directed} source {sources[i]} batch {k}")
        assert sweeps > 0


@pytest.mark.parametrize("group", ["1", "2", "16"])
def test_many_sources_lane_groups(group, monkeypatch):
    """33 sources -> rows of 40 = 5 pieces of 8: with DPPR_PULL_GROUP=16 eight adjacent lanes share a vertex (one chunk
    group, 3 idle lanes of 8); with 2 two lanes share it (3 chunk groups); with 1 every lane has its own vertex (5 chunk
    groups).  Same answers every way."""
    force_dense(monkeypatch, div="1e15", tiers=(4, 120, 16))
    monkeypatch.setenv("DPPR_PULL_GROUP", group)
    V, M, directed = 2_500, 30_000, True
    edges = graphgen.rmat_directed(V, M, seed=11)
    wl = stream.workload(M, 0.1, 0, 0.03, 3)
    sources = [int(x) for x in graphgen.top_out_degree(V, edges, directed, 20)] + list(range(1, 14))
    eps = 1e-8
    oracles = []
    for s in sources:
        o = orc.Oracle(V, directed, edges, wl.W, wl.B, s, eps, 0)
        o.initial_solve()
        oracles.append(o)
    sweeps = 0
    with DynamicPPR(V, directed, wl.W, wl.B, sources, epsilon=eps) as eng:
        eng.init_window_pairs(edges[: wl.W])
        eng.solve_initial()
        for k in range(wl.n_batches + 1):
            if k > 0:
                lo = wl.W + (k - 1) * wl.B
                eng.slide_pairs(edges[lo: lo + wl.B])
                for o in oracles:
                    o.slide(wl.B)
            st = eng.stats()
            assert st.error_flags == 0
            sweeps += st.dense_sweeps
            for i, o in enumerate(oracles):
                check_against(eng.estimates(i), eng.residuals(i), o.p, None, eps, f"group {group} source {sources[i]} batch {k}")
    assert sweeps > 0


@pytest.mark.parametrize("variant,panel,dense", [(0, 8, True), (0, 5, False), (2, 16, True), (1, 1, False), (3, 12, True)])
def test_source_panels_match_single_source_oracles(variant, panel, dense, monkeypatch):
    """33 sources refreshed in panels of at most `panel` (tuning.panel_sources): 5 panels of 7, 7 of 5, 3 of 11, 33 of 1, 3 of
    11 -- every panel its own [V][Sr] block and its own launches.  Same answers as one oracle per source; top-k, validation,
    state export / import and the batch counters go through the panel addressing too."""
    if dense:
        force_dense(monkeypatch, div="1e15", tiers=(4, 120, 16))
    V, M, directed = 2_500, 30_000, variant != 2
    edges = graphgen.rmat_directed(V, M, seed=11) if directed else graphgen.powerlaw_undirected(V, M, seed=12)
    wl = stream.workload(M, 0.1, 0, 0.03, 3)
    sources = [int(x) for x in graphgen.top_out_degree(V, edges, directed, 20)] + list(range(1, 14))
    eps = 1e-8
    oracles = []
    for s in sources:
        o = orc.Oracle(V, directed, edges, wl.W, wl.B, s, eps, variant)
        o.initial_solve()
        oracles.append(o)
    sweeps = 0
    with DynamicPPR(V, directed, wl.W, wl.B, sources, epsilon=eps, variant=variant,
                    tuning={"panel_sources": panel, "dense": 1 if dense else -1}) as eng:
        eng.init_window_pairs(edges[: wl.W])
        eng.solve_initial()
        for k in range(wl.n_batches + 1):
            if k > 0:
                lo = wl.W + (k - 1) * wl.B
                eng.slide_pairs(edges[lo: lo + wl.B])
                for o in oracles:
                    o.slide(wl.B)
            st = eng.stats()
            assert st.error_flags == 0 and st.frontier_pops > 0
            sweeps += st.dense_sweeps
            for i, o in enumerate(oracles):
                check_against(eng.estimates(i), eng.residuals(i), o.p, None, eps, f"panel {panel} source {sources[i]} batch {k}")
        assert (sweeps > 0) == dense
        ids, vals = eng.topk(10)
        sub_ids, sub_vals = eng.topk(10, first_source=6, n_sources=20)   # a range that starts and ends inside panels
        for i in range(len(sources)):
            p = eng.estimates(i)
            order = np.lexsort((np.arange(V), -p))[:10]
            np.testing.assert_array_equal(ids[i], order)
            np.testing.assert_array_equal(vals[i], p[order])
            if 6 <= i < 26:
                np.testing.assert_array_equal(sub_ids[i - 6], order)
            max_r, defect = eng.validate(i)
            assert max_r <= eps and defect <= 1e-12
        # state import lands in the right panel and column
        i = len(sources) - 2
        p, r, neighbour = eng.estimates(i), eng.residuals(i), eng.estimates(i - 1)
        p2 = p.copy(); p2[17] += 0.25
        eng.set_state(i, p2, r)
        np.testing.assert_array_equal(eng.estimates(i), p2)
        np.testing.assert_array_equal(eng.estimates(i - 1), neighbour)
        np.testing.assert_array_equal(eng.topk(1, first_source=i, n_sources=1)[0][0], [17])


def test_dense_off_keeps_no_out_lists(monkeypatch):
    monkeypatch.setenv("DPPR_DENSE_DIV", "0")
    V, M = 3_000, 30_000
    edges = graphgen.rmat_directed(V, M, seed=9)
    wl = stream.workload(M, 0.1, 0, 0.05, 3)
    with DynamicPPR(V, True, wl.W, wl.B, [1]) as eng:
        eng.init_window_pairs(edges[: wl.W])
        eng.solve_initial()
        eng.slide_pairs(edges[wl.W: wl.W + wl.B])
        assert eng.export_window_out_csr() is None
        assert eng.stats().dense_sweeps == 0
