"""Host-side restatement of the arithmetic behind `pacc` (csrc/pull.cuh, pull_finish_unit / pull_fold): during a sweep
episode the amounts popped from a (vertex, source) pair -- bf16 values -- are summed in FP32, every addition is checked
with TwoSum, and an addition that would round sends its amount to the FP64 estimate directly.  The claim the kernel
relies on: whatever the sequence, (FP32 sum) + (amounts sent directly) is EXACTLY the sum of all amounts.  numpy float32
has the same round-to-nearest-even arithmetic as the device (no FMA is involved: the kernel only adds and subtracts)."""
import numpy as np
import pytest


def bf16_trunc(x):
    """float64 -> the bf16 value obtained by truncation towards zero (pull.cuh, bf16_trunc), as float32"""
    f = np.asarray(x, dtype=np.float64).astype(np.float32)          # (round to nearest here; the low bits are cut next)
    f = np.where(np.abs(f.astype(np.float64)) > np.abs(x), np.nextafter(f, np.float32(0)), f).astype(np.float32)  # towards zero
    bits = f.view(np.uint32) & np.uint32(0xffff0000)
    return bits.view(np.float32)


def accumulate(amounts):
    """the kernel's loop: returns (FP32 sum, list of amounts that went to p directly)"""
    pa = np.float32(0.0)
    direct = []
    for av in amounts:
        av = np.float32(av)
        t = np.float32(pa + av)
        bb = np.float32(t - pa)
        err = np.float32(np.float32(pa - np.float32(t - bb)) + np.float32(av - bb))
        if err == np.float32(0.0):
            pa = t
        else:
            direct.append(av)
    return pa, direct


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("shape", ["geometric", "wide", "mixed_sign", "cancelling"])
def test_fp32_amount_sum_plus_direct_amounts_is_exact(seed, shape):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(2, 48))                       # sweeps of an episode
    if shape == "geometric":                           # what an episode looks like: amounts shrinking by ~0.56-0.85 per sweep
        x = 1e-4 * np.cumprod(rng.uniform(0.5, 0.9, n))
    elif shape == "wide":                              # a source vertex: 0.1 down to the tolerance, 8 decades
        x = 10.0 ** rng.uniform(-9.5, -1.0, n)
    elif shape == "mixed_sign":                        # Chebyshev-accelerated amounts change sign
        x = 1e-5 * np.cumprod(rng.uniform(0.4, 0.9, n)) * rng.choice([-1.0, 1.0], n)
    else:                                              # near-total cancellation, then small amounts
        big = 10.0 ** rng.uniform(-3, -2)
        x = np.concatenate([[big, -big * (1 - 2.0 ** -7)], 10.0 ** rng.uniform(-9, -7, n)])
    amounts = bf16_trunc(x)
    assert np.all(np.abs(amounts.astype(np.float64)) <= np.abs(x))          # truncation never pops more than there is
    pa, direct = accumulate(amounts)
    # exact reference: bf16 values are 8-bit significands between 2^-34 and 2^0 here, < 64 of them: float64 adds them exactly
    exact = float(np.sum(amounts.astype(np.float64)))
    got = float(np.float64(pa) + np.sum(np.asarray(direct, dtype=np.float64)))
    assert got == exact
    if shape == "geometric" and n <= 12:               # (0.5^12 = 2^-12: every partial sum still fits 24 bits)
        assert not direct, "amounts within 2^12 of each other never leave the FP32 sum"


def test_an_addition_that_would_round_is_detected():
    big, tiny = bf16_trunc(0.1), bf16_trunc(3e-10)      # 2^28 apart: 0.1 + 3e-10 is not an FP32 number
    pa, direct = accumulate([big, tiny])
    assert pa == big and direct == [tiny]
    pa, direct = accumulate([tiny, big])                 # order does not matter for the check
    assert float(np.float64(pa) + sum(map(float, direct))) == float(big) + float(tiny)
