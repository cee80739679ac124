"""ctypes binding of include/dppr.h (one-to-one; no logic lives here)."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OPTIMIZED, FAST_FRONTIER, EAGER, VANILLA = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_STEPWISE, ENGINE_LEVELSYNC = 0, 1, 3

# every symbol include/dppr.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = [
    "dppr_version", "dppr_last_error", "dppr_create", "dppr_destroy", "dppr_init_window",
    "dppr_init_window_pairs", "dppr_init_window_device_pairs", "dppr_generate_rmat_device", "dppr_generate_stream_device",
    "dppr_generate_stream_host", "dppr_rank_by_degree", "dppr_solve_initial", "dppr_apply_batch", "dppr_apply_batch_pairs",
    "dppr_apply_batch_device_pairs", "dppr_refresh", "dppr_slide", "dppr_slide_pairs",
    "dppr_slide_device_pairs", "dppr_sync", "dppr_get_batch_stats", "dppr_batches_done",
    "dppr_get_estimates", "dppr_get_residuals", "dppr_copy_estimates_device", "dppr_export_window_csr", "dppr_export_window_out_csr",
    "dppr_window_csr_entries", "dppr_set_state", "dppr_repair_only", "dppr_test_sort_pairs",
    "dppr_test_exclusive_scan", "dppr_test_relabel_slot", "dppr_debug_iterlog", "dppr_debug_ctalog", "dppr_kernel_launches",
    "dppr_wait_event", "dppr_get_topk", "dppr_validate", "dppr_check_window_device", "dppr_check_window",
]


class DpprError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"dppr error {code}: {msg}")
        self.code = code


class Tuning(C.Structure):
    _fields_ = [
        ("relabel", C.c_int32), ("relabel_blocks", C.c_int32), ("relabel_both", C.c_int32), ("ctas_per_sm", C.c_int32),
        ("tile_cap", C.c_int32), ("max_iters", C.c_int32), ("dense", C.c_int32), ("pull_group", C.c_int32),
        ("pull_warp_min", C.c_int32), ("pull_big_min", C.c_int32), ("pull_big_chunk", C.c_int32), ("window_path", C.c_int32),
        ("iterlog", C.c_int32), ("probe_iter", C.c_int32), ("dense_div", C.c_double), ("dense_min_edges", C.c_double),
        ("carry_gamma", C.c_double), ("carry_scale", C.c_double), ("dense_accel", C.c_int32), ("signed_push", C.c_int32), ("panel_sources", C.c_int32),
        ("pull_warp_units", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class Config(C.Structure):
    _fields_ = [
        ("vertex_count", C.c_int32), ("directed", C.c_int32), ("window_edges", C.c_int64),
        ("max_batch_edges", C.c_int64), ("alpha", C.c_double), ("epsilon", C.c_double),
        ("variant", C.c_int32), ("device", C.c_int32), ("n_sources", C.c_int32),
        ("sources", C.POINTER(C.c_int32)), ("engine_mode", C.c_int32), ("record_timing", C.c_int32),
        ("pool_factor", C.c_double), ("frontier_capacity", C.c_int64), ("hub_degree", C.c_int32),
        ("reserved0", C.c_int32), ("tuning", Tuning),
    ]


class BatchStats(C.Structure):
    _fields_ = [
        ("batch_index", C.c_int64), ("edges", C.c_int64), ("batch_entries", C.c_int64),
        ("touched_vertices", C.c_int64), ("iterations", C.c_int64), ("frontier_pops", C.c_int64),
        ("traversed_edges", C.c_int64), ("hub_pops", C.c_int64), ("relocations", C.c_int64),
        ("pool_used", C.c_int64), ("ms_upload", C.c_float), ("ms_window", C.c_float),
        ("ms_repair", C.c_float), ("ms_push", C.c_float), ("error_flags", C.c_int32), ("dense_sweeps", C.c_int32),
        ("scatter_edges", C.c_int64), ("dense_slots", C.c_int64), ("dense_pairs", C.c_int64), ("dense_units", C.c_int64), ("dense_pops", C.c_int64),
        ("pool_leaked", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved0"}


def library_path() -> str:
    # DPPR_LIB: A/B runs of differently tuned builds of the same library (development aid)
    return os.environ.get("DPPR_LIB") or os.path.join(_HERE, "lib", "libdppr.so")


def load_library():
    """Load libdppr.so.  Raises loudly if it is missing: there is no fallback implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build the CUDA library first (`make lib` or `python -c 'import "
            "__graft_entry__ as g; g.build()'`).  dynamicppr_b200 has no CPU fallback.")
    L = C.CDLL(path)
    i32p, f64p, vp = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_void_p
    L.dppr_version.restype = C.c_int
    L.dppr_last_error.argtypes = [vp]; L.dppr_last_error.restype = C.c_char_p
    L.dppr_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.dppr_destroy.argtypes = [vp]; L.dppr_destroy.restype = None
    L.dppr_init_window.argtypes = [vp, i32p, i32p, C.c_int64]
    L.dppr_init_window_pairs.argtypes = [vp, i32p, C.c_int64]
    L.dppr_init_window_device_pairs.argtypes = [vp, vp, C.c_int64]
    L.dppr_generate_rmat_device.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_uint64, vp]
    L.dppr_generate_stream_device.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_uint64, vp]
    L.dppr_generate_stream_host.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_uint64, vp, C.c_int32]
    L.dppr_rank_by_degree.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int64, C.c_int32, i32p, i32p, i32p]
    L.dppr_solve_initial.argtypes = [vp]
    for name in ("dppr_apply_batch", "dppr_slide"):
        getattr(L, name).argtypes = [vp, i32p, i32p, C.c_int64]
    for name in ("dppr_apply_batch_pairs", "dppr_slide_pairs"):
        getattr(L, name).argtypes = [vp, i32p, C.c_int64]
    for name in ("dppr_apply_batch_device_pairs", "dppr_slide_device_pairs"):
        getattr(L, name).argtypes = [vp, vp, C.c_int64]
    L.dppr_refresh.argtypes = [vp]
    L.dppr_sync.argtypes = [vp]
    L.dppr_wait_event.argtypes = [vp, vp]
    L.dppr_get_topk.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, i32p, f64p]
    L.dppr_validate.argtypes = [vp, C.c_int32, f64p, f64p]
    L.dppr_check_window_device.argtypes = [vp, vp, C.c_int64, C.POINTER(C.c_int64)]
    L.dppr_check_window.argtypes = [vp, i32p, C.c_int64, C.POINTER(C.c_int64)]
    L.dppr_get_batch_stats.argtypes = [vp, C.c_int64, C.POINTER(BatchStats)]
    L.dppr_batches_done.argtypes = [vp]; L.dppr_batches_done.restype = C.c_int64
    L.dppr_get_estimates.argtypes = [vp, C.c_int32, f64p]
    L.dppr_get_residuals.argtypes = [vp, C.c_int32, f64p]
    L.dppr_copy_estimates_device.argtypes = [vp, C.c_int32, vp]
    L.dppr_export_window_csr.argtypes = [vp, i32p, i32p, i32p]
    L.dppr_export_window_out_csr.argtypes = [vp, i32p, i32p]
    L.dppr_window_csr_entries.argtypes = [vp]; L.dppr_window_csr_entries.restype = C.c_int64
    L.dppr_set_state.argtypes = [vp, C.c_int32, f64p, f64p]
    L.dppr_repair_only.argtypes = [vp]
    L.dppr_kernel_launches.restype = C.c_uint64
    L.dppr_debug_ctalog.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int32, C.POINTER(C.c_int32)]
    L.dppr_debug_iterlog.argtypes = [vp, C.POINTER(C.c_uint32), C.c_int32, C.POINTER(C.c_int32)]
    L.dppr_test_sort_pairs.argtypes = [C.c_int32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_int64, C.c_int32]
    L.dppr_test_exclusive_scan.argtypes = [C.c_int32, C.POINTER(C.c_uint32), C.c_int64, C.POINTER(C.c_uint64)]
    L.dppr_test_relabel_slot.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    L.dppr_test_relabel_slot.restype = C.c_uint32
    _LIB = L
    return L


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class DynamicPPR:
    """One engine = one GPU.  Mirrors the C ABI call for call."""

    def __init__(self, vertex_count, directed, window_edges, max_batch_edges, sources, epsilon=1e-9, variant=0,
                 device=0, engine_mode=ENGINE_AUTO, record_timing=True, alpha=0.15, pool_factor=0.0,
                 frontier_capacity=0, hub_degree=0, tuning=None):
        self.L = load_library()
        self.V = int(vertex_count)
        self._sources = np.ascontiguousarray(np.atleast_1d(sources), dtype=np.int32)
        cfg = Config(vertex_count=self.V, directed=int(bool(directed)), window_edges=int(window_edges),
                     max_batch_edges=int(max_batch_edges), alpha=alpha, epsilon=epsilon, variant=int(variant),
                     device=int(device), n_sources=len(self._sources), sources=_i32(self._sources),
                     engine_mode=int(engine_mode), record_timing=int(bool(record_timing)), pool_factor=pool_factor,
                     frontier_capacity=int(frontier_capacity), hub_degree=int(hub_degree), reserved0=0)
        for name, value in (tuning or {}).items():  # dppr_tuning fields by name (include/dppr.h); 0 = default
            if name not in dict(Tuning._fields_):
                raise ValueError(f"unknown tuning field {name!r}")
            setattr(cfg.tuning, name, value)
        self.h = C.c_void_p()
        rc = self.L.dppr_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            msg = self.L.dppr_last_error(None).decode()
            self.h = None
            raise DpprError(rc, msg)
        self.n_sources = len(self._sources)

    # -- plumbing --------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise DpprError(rc, self.L.dppr_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.dppr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def _pairs(edges):
        e = np.ascontiguousarray(edges, dtype=np.int32)
        assert e.ndim == 2 and e.shape[1] == 2
        return e

    # -- the C ABI ---------------------------------------------------------------------------------
    def init_window_pairs(self, edges):
        e = self._pairs(edges)
        self._check(self.L.dppr_init_window_pairs(self.h, _i32(e), len(e)))

    def init_window(self, edge1, edge2):
        a = np.ascontiguousarray(edge1, np.int32); b = np.ascontiguousarray(edge2, np.int32)
        self._check(self.L.dppr_init_window(self.h, _i32(a), _i32(b), len(a)))

    def init_window_device_pairs(self, device_ptr, n):
        self._check(self.L.dppr_init_window_device_pairs(self.h, C.c_void_p(int(device_ptr)), int(n)))

    def solve_initial(self):
        self._check(self.L.dppr_solve_initial(self.h))

    def apply_batch_pairs(self, edges):
        e = self._pairs(edges)
        self._check(self.L.dppr_apply_batch_pairs(self.h, _i32(e), len(e)))

    def apply_batch(self, edge1, edge2):
        a = np.ascontiguousarray(edge1, np.int32); b = np.ascontiguousarray(edge2, np.int32)
        self._check(self.L.dppr_apply_batch(self.h, _i32(a), _i32(b), len(a)))

    def refresh(self):
        self._check(self.L.dppr_refresh(self.h))

    def repair_only(self):
        self._check(self.L.dppr_repair_only(self.h))

    def slide_pairs(self, edges):
        e = self._pairs(edges)
        self._check(self.L.dppr_slide_pairs(self.h, _i32(e), len(e)))

    def slide(self, edge1, edge2):
        a = np.ascontiguousarray(edge1, np.int32); b = np.ascontiguousarray(edge2, np.int32)
        self._check(self.L.dppr_slide(self.h, _i32(a), _i32(b), len(a)))

    def slide_device_pairs(self, device_ptr, B):
        self._check(self.L.dppr_slide_device_pairs(self.h, C.c_void_p(int(device_ptr)), int(B)))

    def apply_batch_device_pairs(self, device_ptr, B):
        self._check(self.L.dppr_apply_batch_device_pairs(self.h, C.c_void_p(int(device_ptr)), int(B)))

    def sync(self):
        self._check(self.L.dppr_sync(self.h))

    def wait_event(self, cuda_event):
        self._check(self.L.dppr_wait_event(self.h, C.c_void_p(int(cuda_event))))

    def topk(self, k, first_source=0, n_sources=None):
        """(ids[n, k], values[n, k]): the k largest estimates per source, selected on the device"""
        n = self.n_sources - first_source if n_sources is None else int(n_sources)
        ids = np.empty((n, k), np.int32); vals = np.empty((n, k), np.float64)
        self._check(self.L.dppr_get_topk(self.h, first_source, n, k, _i32(ids), vals.ctypes.data_as(C.POINTER(C.c_double))))
        return ids, vals

    def validate(self, source_index=0, invariant=True):
        """(max |r|, largest push-invariant defect) computed on the device (the reference's -DVALIDATE checks)"""
        a, b = C.c_double(0.0), C.c_double(0.0)
        self._check(self.L.dppr_validate(self.h, source_index, C.byref(a), C.byref(b) if invariant else None))
        return a.value, b.value

    def check_window(self, edges):
        """same as check_window_device for W window edges in host memory"""
        e = self._pairs(edges)
        bad = C.c_int64(-1)
        self._check(self.L.dppr_check_window(self.h, _i32(e), len(e), C.byref(bad)))
        return int(bad.value)

    def check_window_device(self, device_ptr, n):
        """entries of the canonical window graph that differ from the one built from these W device-resident edges"""
        bad = C.c_int64(-1)
        self._check(self.L.dppr_check_window_device(self.h, C.c_void_p(int(device_ptr)), int(n), C.byref(bad)))
        return int(bad.value)

    def stats(self, batch_index=-1) -> BatchStats:
        s = BatchStats()
        self._check(self.L.dppr_get_batch_stats(self.h, batch_index, C.byref(s)))
        return s

    def batches_done(self):
        return int(self.L.dppr_batches_done(self.h))

    def estimates(self, source_index=0):
        out = np.empty(self.V, np.float64)
        self._check(self.L.dppr_get_estimates(self.h, source_index, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def residuals(self, source_index=0):
        out = np.empty(self.V, np.float64)
        self._check(self.L.dppr_get_residuals(self.h, source_index, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def copy_estimates_device(self, source_index, device_ptr):
        self._check(self.L.dppr_copy_estimates_device(self.h, source_index, C.c_void_p(int(device_ptr))))

    def window_csr_entries(self):
        return int(self.L.dppr_window_csr_entries(self.h))

    def export_window_csr(self):
        E = self.window_csr_entries()
        rp = np.empty(self.V + 1, np.int32); ci = np.empty(max(E, 1), np.int32); od = np.empty(self.V, np.int32)
        self._check(self.L.dppr_export_window_csr(self.h, _i32(rp), _i32(ci), _i32(od)))
        return rp, ci[:E], od

    def export_window_out_csr(self):
        """test hook: (row_ptr, col_ind) of the out-lists, or None when the engine keeps none"""
        E = self.window_csr_entries()
        rp = np.empty(self.V + 1, np.int32); ci = np.empty(max(E, 1), np.int32)
        rc = self.L.dppr_export_window_out_csr(self.h, _i32(rp), _i32(ci))
        if rc == 1:
            return None
        self._check(rc)
        return rp, ci[:E]

    def iterlog(self, cap=4096):
        """debug (needs DPPR_ITERLOG=1 at construction): rows (frontier, hub_chunks, t_ns) per push iteration"""
        buf = np.zeros((cap, 4), np.uint32)
        n = C.c_int32(0)
        self._check(self.L.dppr_debug_iterlog(self.h, buf.ctypes.data_as(C.POINTER(C.c_uint32)), cap, C.byref(n)))
        b = buf[: n.value].astype(np.uint64)
        return np.stack([b[:, 0], b[:, 1], b[:, 2] | (b[:, 3] << np.uint64(32))], axis=1)

    def ctalog(self, cap=148 * 16):
        buf = np.zeros((cap, 8), np.uint64)
        n = C.c_int32(0)
        self._check(self.L.dppr_debug_ctalog(self.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), cap, C.byref(n)))
        return buf[: n.value]

    def set_state(self, source_index, p=None, r=None):
        f64p = C.POINTER(C.c_double)
        pp = np.ascontiguousarray(p, np.float64) if p is not None else None
        rr = np.ascontiguousarray(r, np.float64) if r is not None else None
        self._check(self.L.dppr_set_state(self.h, source_index,
                                          pp.ctypes.data_as(f64p) if pp is not None else None,
                                          rr.ctypes.data_as(f64p) if rr is not None else None))


def generate_rmat_device(V, M, seed, device_ptr, device=0):
    """fill M int32 pairs of device memory with a seeded R-MAT stream (see include/dppr.h)"""
    L = load_library()
    rc = L.dppr_generate_rmat_device(device, int(V), int(M), int(seed), C.c_void_p(int(device_ptr)))
    if rc != 0:
        raise DpprError(rc, L.dppr_last_error(None).decode())


STREAM_RMAT, STREAM_POWERLAW = 0, 1


def generate_stream_device(kind, V, first_edge, n_edges, seed, device_ptr, device=0):
    """fill n_edges int32 pairs of DEVICE memory with edges [first_edge, first_edge + n_edges) of a seeded stream"""
    L = load_library()
    rc = L.dppr_generate_stream_device(device, int(kind), int(V), int(first_edge), int(n_edges), int(seed), C.c_void_p(int(device_ptr)))
    if rc != 0:
        raise DpprError(rc, L.dppr_last_error(None).decode())


def generate_stream_host(kind, V, first_edge, n_edges, seed, out=None, threads=0):
    """the host twin of generate_stream_device (bit-identical; no GPU needed).  Returns an (n_edges, 2) int32 array;
    `out` may be a pre-allocated C-contiguous array or memmap of that shape."""
    L = load_library()
    if out is None:
        out = np.empty((int(n_edges), 2), np.int32)
    assert out.dtype == np.int32 and out.shape == (int(n_edges), 2) and out.flags["C_CONTIGUOUS"]
    rc = L.dppr_generate_stream_host(int(kind), int(V), int(first_edge), int(n_edges), int(seed),
                                     C.c_void_p(out.ctypes.data), int(threads) if threads else (os.cpu_count() or 1))
    if rc != 0:
        raise DpprError(rc, L.dppr_last_error(None).decode())
    return out


def rank_by_degree(V, directed, pairs=None, n_edges=None, device_ptr=None, by_out_degree=True, device=0, want_degrees=False):
    """exact degree ranking of a whole stream on the device (include/dppr.h, dppr_rank_by_degree)"""
    L = load_library()
    order = np.empty(int(V), np.int32)
    od = np.empty(int(V), np.int32) if want_degrees else None
    idg = np.empty(int(V), np.int32) if want_degrees else None
    if device_ptr is not None:
        ptr, n, on_dev = C.c_void_p(int(device_ptr)), int(n_edges), 1
    else:
        pairs = np.ascontiguousarray(pairs, np.int32)
        ptr, n, on_dev = C.c_void_p(pairs.ctypes.data), len(pairs), 0
    rc = L.dppr_rank_by_degree(device, int(V), int(bool(directed)), int(bool(by_out_degree)), ptr, n, on_dev, _i32(order),
                               _i32(od) if od is not None else None, _i32(idg) if idg is not None else None)
    if rc != 0:
        raise DpprError(rc, L.dppr_last_error(None).decode())
    return (order, od, idg) if want_degrees else order


def kernel_launches() -> int:
    return int(load_library().dppr_kernel_launches())


def test_sort_pairs(keys, vals, key_bits, device=0):
    L = load_library()
    k = np.ascontiguousarray(keys, np.uint32).copy(); v = np.ascontiguousarray(vals, np.uint32).copy()
    rc = L.dppr_test_sort_pairs(device, k.ctypes.data_as(C.POINTER(C.c_uint32)), v.ctypes.data_as(C.POINTER(C.c_uint32)),
                                len(k), key_bits)
    if rc != 0:
        raise DpprError(rc, L.dppr_last_error(None).decode())
    return k, v


def test_exclusive_scan(data, device=0):
    L = load_library()
    d = np.ascontiguousarray(data, np.uint32).copy()
    tot = C.c_uint64()
    rc = L.dppr_test_exclusive_scan(device, d.ctypes.data_as(C.POINTER(C.c_uint32)), len(d), C.byref(tot))
    if rc != 0:
        raise DpprError(rc, L.dppr_last_error(None).decode())
    return d, int(tot.value)
