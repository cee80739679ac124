"""Window / batch arithmetic of the reference stream reader, for the Python-side drivers.

Mirrors ``SlidingGraphVec::PrepareSlidingGraph`` (reference ``SlidingGraphVec.h:47-66``) including
its truncations; the C++ CLI host (``host/EdgeStream.h``) carries the same code.  tests/ check
both against the oracle restatement and the reference's golden dumps.
"""
from __future__ import annotations
from dataclasses import dataclass


@dataclass
class Workload:
    W: int          # sliding_window_size
    B: int          # gStreamUpdateCountPerBatch
    n_batches: int  # gStreamBatchCount
    total: int      # gStreamUpdateCountTotal (capped at M - W)

    def runnable_batches(self, M: int) -> int:
        """the loop also stops when fewer than B edges remain (SlidingGraphVec.h:221)"""
        return min(self.n_batches, (M - self.W) // self.B) if self.B > 0 else 0


def _trunc_i32(x: float) -> int:
    return int(x)  # C++ double -> int conversion truncates toward zero


def workload(M: int, window_ratio: float = 0.1, mode: int = 0, batch_ratio: float = -1.0, batch_count: int = 0,
             per_batch: int = 0, total: int = 0) -> Workload:
    W = _trunc_i32(float(M) * window_ratio)
    if mode == 0:      # SLIDE_WINDOW_RATIO
        B = int(batch_ratio * W)
        nb = batch_count
        tot = B * nb
    elif mode == 1:    # SLIDE_BATCH_SIZE
        B = per_batch
        tot = total
        nb = (tot + B - 1) // B
    else:
        raise ValueError("gWorkloadConfigType must be 0 or 1 (Arguments.h:51-59)")
    tot = min(tot, M - W)
    return Workload(W, B, nb, tot)
