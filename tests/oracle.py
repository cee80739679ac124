"""ctypes wrapper over oracle/_build/liboracle.so (TEST INFRASTRUCTURE: the checker).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    L = C.CDLL(LIB_PATH)
    i32p, f64p, u8p = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    i64p = C.POINTER(C.c_int64)
    L.dppr_oracle_workload.argtypes = [C.c_int64, C.c_double, C.c_int, C.c_double, C.c_int64, C.c_int64, C.c_int64,
                                       i64p, i64p, i64p, i64p]
    L.dppr_oracle_workload.restype = None
    L.dppr_oracle_create.argtypes = [C.c_int32, C.c_int, i32p, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                     C.c_double, C.c_int]
    L.dppr_oracle_create.restype = C.c_void_p
    L.dppr_oracle_destroy.argtypes = [C.c_void_p]
    L.dppr_oracle_initial_solve.argtypes = [C.c_void_p]
    L.dppr_oracle_set_compat_d2.argtypes = [C.c_void_p, C.c_int]
    L.dppr_oracle_slide.argtypes = [C.c_void_p, C.c_int64]
    L.dppr_oracle_slide.restype = C.c_int
    L.dppr_oracle_p.argtypes = [C.c_void_p]; L.dppr_oracle_p.restype = f64p
    L.dppr_oracle_r.argtypes = [C.c_void_p]; L.dppr_oracle_r.restype = f64p
    L.dppr_oracle_outdeg.argtypes = [C.c_void_p]; L.dppr_oracle_outdeg.restype = i32p
    L.dppr_oracle_iteration_id.argtypes = [C.c_void_p]; L.dppr_oracle_iteration_id.restype = C.c_int32
    L.dppr_oracle_pos.argtypes = [C.c_void_p]; L.dppr_oracle_pos.restype = C.c_int64
    L.dppr_oracle_counters.argtypes = [C.c_void_p, i64p, i64p, i64p]
    L.dppr_oracle_canonical_csr.argtypes = [C.c_void_p, i32p, i32p, i32p]
    L.dppr_oracle_power_iteration.argtypes = [C.c_void_p, f64p]
    L.dppr_oracle_power_iteration.restype = C.c_int
    L.dppr_oracle_repair_sequential.argtypes = [C.c_int64, i32p, i32p, u8p, C.c_int32, f64p, f64p, i32p]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def workload(M, window_ratio=0.1, mode=0, batch_ratio=0.01, batch_count=100, per_batch=0, total=0):
    """(W, B, n_batches, total) by the reference's arithmetic (SlidingGraphVec.h:47-66)."""
    L = load()
    out = [C.c_int64() for _ in range(4)]
    L.dppr_oracle_workload(M, window_ratio, mode, batch_ratio, batch_count, per_batch, total,
                           *[C.byref(x) for x in out])
    return tuple(int(x.value) for x in out)


def repair_sequential(u, v, is_insert, source, p, r, deg):
    """In-place sequential residual repair; deg holds PRE-batch out-degrees."""
    L = load()
    u = np.ascontiguousarray(u, np.int32); v = np.ascontiguousarray(v, np.int32)
    ins = np.ascontiguousarray(is_insert, np.uint8)
    assert r.dtype == np.float64 and deg.dtype == np.int32 and p.dtype == np.float64
    L.dppr_oracle_repair_sequential(len(u), _p(u, C.c_int32), _p(v, C.c_int32), _p(ins, C.c_uint8), source,
                                    _p(p, C.c_double), _p(r, C.c_double), _p(deg, C.c_int32))


class Oracle:
    """Single-source streaming reverse-push PPR, CPU restatement of the reference."""

    def __init__(self, V, directed, edges, W, max_batch, source, eps=1e-9, variant=0, compat_d2=False):
        self.L = load()
        self.edges = np.ascontiguousarray(edges, dtype=np.int32)  # keep alive: borrowed by C
        self.V, self.directed, self.W = int(V), bool(directed), int(W)
        self.Ew = self.W if directed else 2 * self.W
        self.h = self.L.dppr_oracle_create(V, int(directed), _p(self.edges, C.c_int32), len(self.edges), W,
                                           max_batch, source, eps, variant)
        if not self.h:
            raise ValueError("dppr_oracle_create rejected the arguments")
        # reference defect D2 (seed-dedupe sentinel collision); golden pinning only
        self.L.dppr_oracle_set_compat_d2(self.h, int(compat_d2))

    def close(self):
        if self.h:
            self.L.dppr_oracle_destroy(self.h)
            self.h = None

    __del__ = close

    def initial_solve(self):
        self.L.dppr_oracle_initial_solve(self.h)

    def slide(self, B):
        return self.L.dppr_oracle_slide(self.h, B)

    @property
    def p(self):
        return np.ctypeslib.as_array(self.L.dppr_oracle_p(self.h), shape=(self.V,)).copy()

    @property
    def r(self):
        return np.ctypeslib.as_array(self.L.dppr_oracle_r(self.h), shape=(self.V,)).copy()

    @property
    def outdeg(self):
        return np.ctypeslib.as_array(self.L.dppr_oracle_outdeg(self.h), shape=(self.V,)).copy()

    @property
    def iteration_id(self):
        return int(self.L.dppr_oracle_iteration_id(self.h))

    def counters(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.L.dppr_oracle_counters(self.h, C.byref(a), C.byref(b), C.byref(c))
        return dict(iterations=a.value, pops=b.value, traversed=c.value)

    def canonical_csr(self):
        rp = np.empty(self.V + 1, np.int32); ci = np.empty(max(self.Ew, 1), np.int32); od = np.empty(self.V, np.int32)
        self.L.dppr_oracle_canonical_csr(self.h, _p(rp, C.c_int32), _p(ci, C.c_int32), _p(od, C.c_int32))
        return rp, ci[: self.Ew], od

    def power_iteration(self):
        out = np.empty(self.V, np.float64)
        self.L.dppr_oracle_power_iteration(self.h, _p(out, C.c_double))
        return out
