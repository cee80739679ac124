"""Seeded synthetic edge streams in the reference encoder's ``.bin`` format.

Format (writer: reference ``encoder/GraphEncoder.h:86-95``; readers ``GraphVec.h:34-70``,
``SlidingGraphVec.h:35-47``): native little-endian ``int32 vertex_count`` followed by M
records ``(int32 v1, int32 v2)``; no edge count (M = (size-4)/8); record order IS the stream
order; an undirected graph stores each edge once (``-i 0`` mirrors at load).

Two shapes:
  * ``powerlaw_undirected``: Chung-Lu style endpoint sampling from a Zipf weight vector, no
    self-loops, no duplicate undirected pairs, ids randomly permuted, order shuffled.
  * ``rmat_directed``: R-MAT (a,b,c,d = 0.57,0.19,0.19,0.05), scale ceil(log2 V), ids folded
    to < V; duplicates and self-loops are kept (the reference keeps multi-edges).

Everything is driven by ``numpy.random.Generator(PCG64(seed))`` so the container that makes
the golden fixtures and the GPU box (same image, same numpy) produce identical files.
"""
from __future__ import annotations

import os
import numpy as np

# BASELINE.json configs -> (V, M, directed)
SHAPES = {
    "dblp": (317_080, 1_049_866, False),
    "youtube": (1_134_890, 2_987_624, False),
    "livejournal": (4_847_571, 68_993_773, True),
    "orkut": (3_072_441, 117_185_083, False),
    "twitter": (41_652_230, 1_468_365_182, True),
}
BASE_SEED = 20261017


def powerlaw_undirected(V: int, M: int, seed: int, exponent: float = 0.75) -> np.ndarray:
    """Return an (M, 2) int32 array of distinct undirected edges (each stored once)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    w = (np.arange(1, V + 1, dtype=np.float64)) ** (-exponent)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    perm = rng.permutation(V).astype(np.int64)
    have = np.empty(0, dtype=np.int64)
    need = M
    while True:
        n = int(need * 1.15) + 1024
        a = np.searchsorted(cdf, rng.random(n), side="right")
        b = np.searchsorted(cdf, rng.random(n), side="right")
        np.minimum(a, V - 1, out=a)
        np.minimum(b, V - 1, out=b)
        keep = a != b
        a, b = a[keep], b[keep]
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        key = lo.astype(np.int64) * V + hi.astype(np.int64)
        have = np.unique(np.concatenate([have, key]))
        if have.size >= M:
            break
        need = M - have.size
    key = rng.permutation(have)[:M]
    lo, hi = key // V, key % V
    # random orientation of the stored pair, random vertex relabelling
    flip = rng.random(M) < 0.5
    e1 = np.where(flip, hi, lo)
    e2 = np.where(flip, lo, hi)
    out = np.empty((M, 2), dtype=np.int32)
    out[:, 0] = perm[e1]
    out[:, 1] = perm[e2]
    return out


def rmat_directed(V: int, M: int, seed: int, abcd=(0.57, 0.19, 0.19, 0.05), chunk: int = 1 << 24) -> np.ndarray:
    """Return an (M, 2) int32 array of directed R-MAT edges, ids folded into [0, V)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    scale = int(np.ceil(np.log2(max(V, 2))))
    a, b, c, _ = abcd
    out = np.empty((M, 2), dtype=np.int32)
    # a fixed relabelling so that low ids are not the hubs
    mult = 2654435761 % V
    while np.gcd(mult, V) != 1:
        mult += 1
    add = int(rng.integers(0, V))
    done = 0
    while done < M:
        n = min(chunk, M - done)
        src = np.zeros(n, dtype=np.int64)
        dst = np.zeros(n, dtype=np.int64)
        for _ in range(scale):
            r = rng.random(n)
            sbit = r >= a + b                      # quadrants c, d -> source bit 1
            dbit = ((r >= a) & (r < a + b)) | (r >= a + b + c)  # quadrants b, d -> dest bit 1
            src = (src << 1) | sbit
            dst = (dst << 1) | dbit
        src %= V
        dst %= V
        out[done:done + n, 0] = (src * mult + add) % V
        out[done:done + n, 1] = (dst * mult + add) % V
        done += n
    return out


def write_bin(path: str, V: int, edges: np.ndarray) -> None:
    """Write the encoder format: int32 V, then (v1, v2) int32 pairs in stream order."""
    edges = np.ascontiguousarray(edges, dtype=np.int32)
    assert edges.ndim == 2 and edges.shape[1] == 2
    assert edges.size == 0 or (edges.min() >= 0 and edges.max() < V)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(np.int32(V).tobytes())
        f.write(edges.tobytes())
    os.replace(tmp, path)


def read_bin(path: str):
    """Return (V, edges[M,2] int32) from an encoder-format file (memory-mapped)."""
    size = os.path.getsize(path)
    assert size >= 4 and (size - 4) % 8 == 0, "not an encoder .bin file"
    V = int(np.fromfile(path, dtype=np.int32, count=1)[0])
    M = (size - 4) // 8
    edges = np.memmap(path, dtype=np.int32, mode="r", offset=4, shape=(M, 2))
    return V, edges


def make_shape(name: str, path: str, scale: float = 1.0, seed: int | None = None):
    """Generate one of the BASELINE.json shapes (optionally scaled down) into ``path``."""
    V, M, directed = SHAPES[name]
    V = max(16, int(V * scale))
    M = max(64, int(M * scale))
    if seed is None:
        seed = BASE_SEED + list(SHAPES).index(name)
    edges = rmat_directed(V, M, seed) if directed else powerlaw_undirected(V, M, seed)
    write_bin(path, V, edges)
    return V, M, directed


def top_out_degree(V: int, edges: np.ndarray, directed: bool, k: int) -> np.ndarray:
    """Exact top-k vertices by full-file out-degree, ties broken by smaller id.

    Mirrors reference ``workload/Graph.h:178-198`` (``ChooseVertexDegreeRange`` with
    ``rank_ed - rank_st == num``: the exact ranks, descending degree); the reference uses
    an unstable ``std::sort`` so tie order there is unspecified -- ours is deterministic.
    """
    deg = np.bincount(edges[:, 0], minlength=V).astype(np.int64)
    if not directed:
        deg += np.bincount(edges[:, 1], minlength=V)
    order = np.lexsort((np.arange(V), -deg))
    return order[:k].astype(np.int32)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("shape", choices=list(SHAPES))
    ap.add_argument("out")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=None)
    a = ap.parse_args()
    V, M, d = make_shape(a.shape, a.out, a.scale, a.seed)
    print(f"wrote {a.out}: V={V} M={M} directed={int(d)}")
