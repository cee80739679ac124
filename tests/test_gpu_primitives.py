"""GPU: the hand-written scan and radix sort (dynamicppr_b200/csrc/primitives.cuh) vs numpy."""
import numpy as np
import pytest

from dynamicppr_b200 import binding

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 255, 256, 2047, 2048, 2049, 4096, 100_003, 2048 * 2048 + 17])
def test_exclusive_scan(n):
    rng = np.random.default_rng(n)
    data = rng.integers(0, 9, size=n, dtype=np.uint32)
    out, total = binding.test_exclusive_scan(data)
    exp = np.concatenate([[0], np.cumsum(data, dtype=np.uint64)[:-1]]).astype(np.uint32) if n else data
    np.testing.assert_array_equal(out, exp)
    assert total == int(data.sum(dtype=np.uint64))


@pytest.mark.parametrize("n,bits", [(0, 8), (1, 8), (2, 1), (33, 5), (2048, 8), (2049, 9), (5000, 16), (70_001, 17),
                                    (1_000_003, 21), (3_000_000, 26), (300_000, 32)])
def test_radix_sort_pairs_is_stable(n, bits):
    rng = np.random.default_rng(n + bits)
    hi = (1 << bits) - 1
    keys = rng.integers(0, hi + 1, size=n, dtype=np.uint64).astype(np.uint32)
    if n > 10:  # plenty of duplicates so that stability is actually exercised
        keys[: n // 2] = keys[n // 2: n // 2 + n // 2] if n % 2 == 0 else keys[: n // 2]
        keys[::3] = keys[0]
    vals = np.arange(n, dtype=np.uint32)
    k, v = binding.test_sort_pairs(keys, vals, bits)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(k, keys[order])
    np.testing.assert_array_equal(v, vals[order])


def test_radix_sort_all_equal_and_presorted():
    n = 10_000
    k, v = binding.test_sort_pairs(np.full(n, 7, np.uint32), np.arange(n, dtype=np.uint32), 8)
    np.testing.assert_array_equal(v, np.arange(n))
    k, v = binding.test_sort_pairs(np.arange(n, dtype=np.uint32)[::-1].copy(), np.arange(n, dtype=np.uint32), 14)
    np.testing.assert_array_equal(k, np.arange(n))
    np.testing.assert_array_equal(v, np.arange(n)[::-1])
