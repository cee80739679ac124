"""Pins oracle/dppr_oracle.c against the REAL reference CPU build (tests/golden/*.npz, written by
tests/golden/make_golden.py from oracle/_ref/ref_harness_serial).  Bit-exact: the restatement
performs the same FP64 operations in the same order as the reference's serial elision."""
import glob
import os
import numpy as np
import pytest

import oracle as orc

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
assert GOLDEN, "golden fixtures missing"


def _flags(s):
    t = s.split()
    return {t[i]: t[i + 1] for i in range(0, len(t), 2)}


def _workload(g):
    f = _flags(str(g["flags"]))
    M = len(g["edges"])
    return orc.workload(M, float(f.get("-w", 0.1)), int(f.get("-n", 0)), float(f.get("-r", -1.0)),
                        int(f.get("-b", 0)), int(f.get("-c", 0)), int(f.get("-l", 0)))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_workload_arithmetic_matches_reference(path):
    g = np.load(path)
    W, B, nb, total = _workload(g)
    assert W == int(g["W"]) and B == int(g["B"])
    # the reference stops when fewer than B edges remain (SlidingGraphVec.h:221)
    M = len(g["edges"])
    expect = min(nb, (M - W) // B)
    assert int(g["n_snap"]) == expect + 1


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_bit_exact_with_reference(path):
    g = np.load(path)
    V, directed, edges = int(g["V"]), bool(g["directed"]), g["edges"]
    W, B, nb, _ = _workload(g)
    for variant in g["variants"]:
        o = orc.Oracle(V, directed, edges, W, B, int(g["source"]), float(g["eps"]), int(variant), compat_d2=True)
        o.initial_solve()
        for k in range(int(g["n_snap"])):
            if k > 0:
                assert o.slide(B) == 0
            tag = f"{os.path.basename(path)} variant {variant} snapshot {k}"
            rp, ci, od = o.canonical_csr()
            np.testing.assert_array_equal(rp, g["in_row_ptr"][k], err_msg=tag)
            np.testing.assert_array_equal(ci, g["in_col"][k], err_msg=tag)
            np.testing.assert_array_equal(od, g["outdeg"][k], err_msg=tag)
            np.testing.assert_array_equal(o.p, g[f"v{variant}_p"][k], err_msg=tag)
            np.testing.assert_array_equal(o.r, g[f"v{variant}_r"][k], err_msg=tag)
            assert o.iteration_id == int(g[f"v{variant}_iteration_id"][k]), tag
        # one more slide must report "stream over" exactly when the reference stopped early
        M = len(edges)
        if int(g["n_snap"]) - 1 < nb:
            assert o.slide(B) == 1
        o.close()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_power_iteration_bit_exact(path):
    g = np.load(path)
    V, directed, edges = int(g["V"]), bool(g["directed"]), g["edges"]
    W, B, _, _ = _workload(g)
    o = orc.Oracle(V, directed, edges, W, B, int(g["source"]), float(g["eps"]), 0)
    o.initial_solve()
    for k in range(min(int(g["n_snap"]), 6)):
        if k > 0:
            o.slide(B)
        np.testing.assert_array_equal(o.power_iteration(), g["pow"][k])
    o.close()


def _d2_possible(g):
    """reference defect D2 needs an offset == V among the 2*batch_length seed slots."""
    blen = 2 * int(g["B"]) * (1 if bool(g["directed"]) else 2)
    return 2 * blen > int(g["V"])


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_reference_within_eps_of_its_power_iteration(path):
    """the margin the 2-eps parity criterion relies on: |p - pi| <= max|r| < eps.
    Holds for the reference wherever defect D2 cannot fire."""
    g = np.load(path)
    if _d2_possible(g):
        pytest.skip("2*batch_length > V: reference defect D2 can double-seed a vertex")
    eps = float(g["eps"])
    for variant in g["variants"]:
        assert np.abs(g[f"v{variant}_r"]).max() < eps
        assert np.abs(g[f"v{variant}_p"] - g["pow"]).max() < eps


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_without_d2_is_identical_or_better(path):
    """The oracle the product is judged against runs with the D2 sentinel fixed.  Where D2 cannot
    fire it must be bit-identical to the reference; where it can, it must satisfy the
    reference's own validator bounds (PPRCPUMTCilkRev.h:293-304) which the reference then misses."""
    g = np.load(path)
    V, directed, edges = int(g["V"]), bool(g["directed"]), g["edges"]
    W, B, _, _ = _workload(g)
    eps = float(g["eps"])
    for variant in g["variants"]:
        o = orc.Oracle(V, directed, edges, W, B, int(g["source"]), eps, int(variant))
        o.initial_solve()
        for k in range(int(g["n_snap"])):
            if k > 0:
                assert o.slide(B) == 0
            if not _d2_possible(g):
                np.testing.assert_array_equal(o.p, g[f"v{variant}_p"][k])
                np.testing.assert_array_equal(o.r, g[f"v{variant}_r"][k])
            else:
                assert np.abs(o.r).max() < eps
                assert np.abs(o.p - g["pow"][k]).max() < eps
        o.close()


def test_unmodified_reference_defect_d2_is_recorded():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dense_multi_directed.npz"))
    assert _d2_possible(g)
    assert np.abs(g["v0_r"]).max() > 1e6 * float(g["eps"])  # a residual the reference leaves behind


def test_unmodified_reference_defect_d1_is_recorded():
    """DESIGN.md 'reference defect D1': the incremental host adjacency of the unmodified reference
    is wrong on undirected streams and right on directed ones."""
    for path in GOLDEN:
        g = np.load(path)
        wrong = [int(g[f"v{v}_ref_inc_rows_differ"]) for v in g["variants"]]
        if bool(g["directed"]):
            assert wrong == [0] * len(wrong), path
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dense_multi_undirected.npz"))
    assert int(g["v0_ref_inc_rows_differ"]) > 0
