"""The BASELINE.json configurations as runnable workloads (bench.py, the full-size GPU tests, scripts/).

Each config names a seeded synthetic stream in the reference encoder's ``.bin`` layout, the reference CLI flags it is
run with (``scripts/gpu.sh:8-18``) and the rule that picks its sources (``workload/Workload.cpp:47-55``).  Streams of
configs 3-5 come from the counter-based generator of ``include/dppr.h`` (``dppr_generate_stream_device`` and its
bit-identical host twin), so the GPU can hold the whole stream in HBM while the host writes only the prefix the
reference CPU implementation reads -- into a SPARSE ``.bin`` file of the full size, because the reference derives
the window size from the file size (``SlidingGraphVec.h:47-48``).  Config 2 keeps round 1's numpy generator.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from . import binding, graphgen, stream

SEED = graphgen.BASE_SEED


@dataclass(frozen=True)
class BaselineConfig:
    index: int            # position in BASELINE.json "configs" (0-based) + 1
    shape: str
    V: int
    M: int
    directed: bool
    kind: int | None      # binding.STREAM_* of the counter-based generator, None = graphgen (numpy)
    mode: int             # -n
    batch_ratio: float    # -r
    batch_count: int      # -b
    per_batch: int        # -c
    total: int            # -l
    n_sources: int        # sources of the whole job
    multi_source: bool    # True: one fixed source list split over the GPUs (strong scaling)
    eps: float = 1e-9
    window_ratio: float = 0.1

    @property
    def seed(self) -> int:
        return SEED + self.index

    def workload(self) -> stream.Workload:
        return stream.workload(self.M, self.window_ratio, self.mode, self.batch_ratio, self.batch_count, self.per_batch, self.total)

    def cli_flags(self, batches: int | None = None):
        """reference CLI flags (Arguments.h:66-86) without -d / -s / -t / -o"""
        f = ["-a", "0", "-i", str(int(self.directed)), "-y", "1", "-w", str(self.window_ratio), "-n", str(self.mode), "-e", repr(self.eps)]
        if self.mode == 0:
            f += ["-r", str(self.batch_ratio), "-b", str(batches if batches is not None else self.batch_count)]
        else:
            f += ["-c", str(self.per_batch), "-l", str((batches if batches is not None else self.workload().n_batches) * self.per_batch)]
        return f

    def describe(self) -> str:
        g = {None: "graphgen.powerlaw_undirected (numpy)", binding.STREAM_RMAT: "device R-MAT generator",
             binding.STREAM_POWERLAW: "device power-law generator"}[self.kind]
        flags = f"-r {self.batch_ratio} (mode 0)" if self.mode == 0 else f"-c {self.per_batch} -l {self.total} (mode 1)"
        return (f"BASELINE configs[{self.index - 1}]: {self.shape}-shaped synthetic {'directed' if self.directed else 'undirected'} stream "
                f"({self.V:,} V, {self.M:,} E; {g}, seed {self.seed}), window {self.window_ratio}, {flags}, eps {self.eps}")


CONFIGS = {
    2: BaselineConfig(2, "youtube", 1_134_890, 2_987_624, False, None, 0, 0.01, 100, 0, 0, 1, False),
    3: BaselineConfig(3, "livejournal", 4_847_571, 68_993_773, True, binding.STREAM_RMAT, 1, -1.0, 0, 100, 10_000, 1, False),
    4: BaselineConfig(4, "orkut", 3_072_441, 117_185_083, False, binding.STREAM_POWERLAW, 0, 0.01, 100, 0, 0, 1000, True),
    5: BaselineConfig(5, "twitter", 41_652_230, 1_468_365_182, True, binding.STREAM_RMAT, 0, 0.01, 100, 0, 0, 64, True),
}


def scaled(cfg: BaselineConfig, scale: float, n_sources: int | None = None) -> BaselineConfig:
    """the same shape at a fraction of the size (tests)"""
    import dataclasses
    return dataclasses.replace(cfg, V=max(64, int(cfg.V * scale)), M=max(1024, int(cfg.M * scale)),
                               n_sources=n_sources if n_sources is not None else cfg.n_sources)


# ---- streams -----------------------------------------------------------------------------------------------------
_numpy_cache: dict = {}


def host_edges(cfg: BaselineConfig, first: int, n: int, out=None) -> np.ndarray:
    """edges [first, first + n) of the config's stream on the host"""
    if cfg.kind is None:
        key = (cfg.shape, cfg.V, cfg.M)
        if key not in _numpy_cache:
            _numpy_cache[key] = graphgen.powerlaw_undirected(cfg.V, cfg.M, SEED + list(graphgen.SHAPES).index(cfg.shape))
        e = _numpy_cache[key][first: first + n]
        if out is not None:
            out[:] = e
            return out
        return e
    return binding.generate_stream_host(cfg.kind, cfg.V, first, n, cfg.seed, out=out)


def device_edges(cfg: BaselineConfig, first: int, n: int, device: int = 0):
    """edges [first, first + n) as an (n, 2) int32 torch tensor on cuda:<device>"""
    import torch
    t = torch.empty((n, 2), dtype=torch.int32, device=f"cuda:{device}")
    if cfg.kind is None:
        t.copy_(torch.from_numpy(np.ascontiguousarray(host_edges(cfg, first, n))))
    else:
        binding.generate_stream_device(cfg.kind, cfg.V, first, n, cfg.seed, t.data_ptr(), device=device)
    return t


def write_prefix_bin(cfg: BaselineConfig, path: str, n_edges: int) -> str:
    """``.bin`` file of the FULL size whose first n_edges records are real and whose tail is a hole (sparse): enough for a
    run that slides fewer than (n_edges - W) / B batches, which never reads further (SlidingGraphVec.h:226-262)."""
    n_edges = min(n_edges, cfg.M)
    full = 4 + 8 * cfg.M
    tag = path + ".ok"
    want = f"{cfg.index} {cfg.V} {cfg.M} {cfg.seed} {n_edges}"
    if os.path.exists(path) and os.path.exists(tag) and os.path.getsize(path) == full:
        have = open(tag).read().split()
        if have[:4] == want.split()[:4] and int(have[4]) >= n_edges:
            return path
    if os.path.exists(tag):
        os.remove(tag)
    with open(path, "wb") as f:
        f.write(np.int32(cfg.V).tobytes())
        f.truncate(full)
    mm = np.memmap(path, dtype=np.int32, mode="r+", offset=4, shape=(n_edges, 2))
    step = 1 << 24
    for lo in range(0, n_edges, step):
        m = min(step, n_edges - lo)
        blk = np.empty((m, 2), np.int32)
        host_edges(cfg, lo, m, out=blk)
        mm[lo: lo + m] = blk
    mm.flush()
    del mm
    with open(tag, "w") as f:
        f.write(want)
    return path


# ---- sources -----------------------------------------------------------------------------------------------------
def host_degree_order(cfg: BaselineConfig) -> np.ndarray:
    """host twin of dppr_rank_by_degree for a config's stream: ids by descending out-degree, ties by ascending id"""
    deg = np.zeros(cfg.V, np.int64)
    step = 1 << 25
    for lo in range(0, cfg.M, step):
        e = host_edges(cfg, lo, min(step, cfg.M - lo))
        deg += np.bincount(e[:, 0], minlength=cfg.V)
        if not cfg.directed:
            deg += np.bincount(e[:, 1], minlength=cfg.V)
    return np.lexsort((np.arange(cfg.V), -deg)).astype(np.int32)


def top_sources(cfg: BaselineConfig, k: int, device: int = 0, dev_stream=None, on_host: bool = False) -> np.ndarray:
    """the exact top-k vertices by out-degree over the WHOLE stream, ties by ascending id (workload/Graph.h:178-198 with
    rank_ed - rank_st == num).  Generated streams are ranked on the device (dppr_rank_by_degree) unless on_host."""
    if cfg.kind is None:
        return graphgen.top_out_degree(cfg.V, host_edges(cfg, 0, cfg.M), cfg.directed, k)
    if on_host:
        return host_degree_order(cfg)[:k].copy()
    import torch
    own = dev_stream is None
    if own:
        dev_stream = device_edges(cfg, 0, cfg.M, device)
    order = binding.rank_by_degree(cfg.V, cfg.directed, n_edges=cfg.M, device_ptr=dev_stream.data_ptr(), device=device)
    if own:
        del dev_stream
        torch.cuda.empty_cache()
    return order[:k].copy()
