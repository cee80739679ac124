"""shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations
import glob
import os
import numpy as np

import oracle as orc
from dynamicppr_b200 import stream

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
GOLDEN_IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


def parse_flags(s):
    t = str(s).split()
    return {t[i]: t[i + 1] for i in range(0, len(t), 2)}


def golden_workload(g) -> stream.Workload:
    f = parse_flags(g["flags"])
    return stream.workload(len(g["edges"]), float(f.get("-w", 0.1)), int(f.get("-n", 0)), float(f.get("-r", -1.0)),
                           int(f.get("-b", 0)), int(f.get("-c", 0)), int(f.get("-l", 0)))


def d2_possible(g) -> bool:
    """reference defect D2: needs an offset == V among the 2*batch_length seed slots."""
    blen = 2 * int(g["B"]) * (1 if bool(g["directed"]) else 2)
    return 2 * blen > int(g["V"])


def have_gpu() -> bool:
    try:
        import ctypes
        cuda = ctypes.CDLL("libcudart.so.12")
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False
    n = ctypes.c_int(0)
    return cuda.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0


def check_against(eng_p, eng_r, ref_p, pow_p, eps, tag):
    """the north-star criterion: within 2 eps of the reference CPU push AND of power iteration;
    residuals within the tolerance (gpu/PPRRevPushGPU.cuh:141-143, non-strict at the boundary)."""
    assert np.all(np.isfinite(eng_p)) and np.all(np.isfinite(eng_r)), tag
    assert np.abs(eng_r).max() <= eps, f"{tag}: max|r| = {np.abs(eng_r).max():.3e} > eps"
    if ref_p is not None:
        d = np.abs(eng_p - ref_p).max()
        assert d <= 2 * eps, f"{tag}: max|p - p_ref| = {d:.3e} > 2 eps"
    if pow_p is not None:
        d = np.abs(eng_p - pow_p).max()
        assert d <= 2 * eps, f"{tag}: max|p - p_pow| = {d:.3e} > 2 eps"
