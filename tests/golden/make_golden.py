#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REAL reference CPU implementation.

Run in the build container only (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

For every case below it writes a seeded synthetic ``.bin`` stream, runs
``oracle/_ref/ref_harness_serial`` (oracle/ref_harness.cpp over the unmodified reference
headers, Cilk serial elision => deterministic) with ``--scratch-graph --pow`` and stores, for
the initial solve and every batch: p, r, out-degrees, canonical in-CSR (rows ascending) and the
reference's own power-iteration vector (PPRCPUPowVec::CalPPRRev).  The edge stream itself is
stored too, so the fixtures are self-contained on the GPU box (no /root/reference there).

``--scratch-graph`` = the reference with the line it keeps commented at
cpu/PPRCPUMTCilk.h:126 enabled; see DESIGN.md "reference defect D1" for why the golden vectors
must not come from the incremental host adjacency on undirected streams.  Each case also
records, from a second run WITHOUT the flag, how many adjacency rows the unmodified reference
got wrong (``ref_inc_rows_differ``) -- zero for directed streams.
"""
from __future__ import annotations
import os
import subprocess
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from dynamicppr_b200 import graphgen  # noqa: E402
from refdump import read_dump  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness_serial")


def dense_multigraph(V, M, seed):
    """uniform random pairs: many duplicate edges and self-loops (kept, like the reference)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, V, size=(M, 2)).astype(np.int32)


def hub_expiry(V, M, seed):
    """a hub whose edges all arrive early (and therefore all expire), vertices that drop to
    degree 0 and re-appear."""
    rng = np.random.Generator(np.random.PCG64(seed))
    e = rng.integers(1, V, size=(M, 2)).astype(np.int32)
    n = M // 8
    e[:n, 1] = 0                      # first n edges all point at vertex 0
    e[M // 2:M // 2 + n // 2, 0] = 0  # later it re-appears as a source
    return e


# name -> (edges builder, V, M, directed, flag string, source chooser)
CASES = {
    "pl_undirected": dict(kind="powerlaw", V=300, M=3000, directed=0, seed=11,
                          flags="-w 0.1 -n 0 -r 0.07 -b 25", source="top", eps=1e-9, variants=[0, 1, 2, 3]),
    "rmat_directed_mode1": dict(kind="rmat", V=512, M=6000, directed=1, seed=12,
                                flags="-w 0.1 -n 1 -c 7 -l 140", source="top", eps=1e-9, variants=[0, 1, 2, 3]),
    "dense_multi_undirected": dict(kind="dense", V=40, M=3000, directed=0, seed=13,
                                   flags="-w 0.1 -n 0 -r 0.13 -b 30", source=1, eps=1e-9, variants=[0, 3]),
    "dense_multi_directed": dict(kind="dense", V=40, M=3000, directed=1, seed=14,
                                 flags="-w 0.2 -n 0 -r 0.11 -b 20", source=1, eps=1e-9, variants=[0, 2]),
    "hub_expiry_directed": dict(kind="hub", V=200, M=4000, directed=1, seed=15,
                                flags="-w 0.25 -n 0 -r 0.1 -b 30", source=0, eps=1e-9, variants=[0, 1]),
    "loose_eps_undirected": dict(kind="powerlaw", V=400, M=4000, directed=0, seed=16,
                                 flags="-w 0.15 -n 0 -r 0.02 -b 40", source="top", eps=1e-5, variants=[0, 1, 2, 3]),
    "batch_of_one": dict(kind="rmat", V=256, M=2000, directed=1, seed=17,
                         flags="-w 0.1 -n 1 -c 1 -l 40", source="top", eps=1e-9, variants=[0]),
    "stream_runs_out": dict(kind="powerlaw", V=200, M=1000, directed=0, seed=18,
                            flags="-w 0.5 -n 0 -r 0.3 -b 10", source="top", eps=1e-9, variants=[0]),
}


def build_edges(c):
    if c["kind"] == "powerlaw":
        return graphgen.powerlaw_undirected(c["V"], c["M"], c["seed"])
    if c["kind"] == "rmat":
        return graphgen.rmat_directed(c["V"], c["M"], c["seed"])
    if c["kind"] == "dense":
        return dense_multigraph(c["V"], c["M"], c["seed"])
    if c["kind"] == "hub":
        return hub_expiry(c["V"], c["M"], c["seed"])
    raise KeyError(c["kind"])


def run_harness(binpath, c, variant, source, dump, scratch):
    cmd = [HARNESS, "-d", binpath, "-a", "0", "-i", str(c["directed"]), "-y", "1", *c["flags"].split(),
           "-s", str(source), "-o", str(variant), "-e", repr(c["eps"]), "--quiet", "--pow", "--dump", dump]
    if scratch:
        cmd.append("--scratch-graph")
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
    kv = {}
    for line in out.splitlines():
        parts = line.split()
        if len(parts) >= 2:
            kv[parts[0]] = parts[1]
    return kv


def main():
    assert os.path.exists(HARNESS), "run `make -C oracle ref` first (needs /root/reference)"
    with tempfile.TemporaryDirectory() as tmp:
        for name, c in CASES.items():
            edges = build_edges(c)
            binpath = os.path.join(tmp, name + ".bin")
            graphgen.write_bin(binpath, c["V"], edges)
            source = c["source"]
            if source == "top":
                source = int(graphgen.top_out_degree(c["V"], edges, bool(c["directed"]), 1)[0])
            out = dict(V=np.int32(c["V"]), directed=np.int32(c["directed"]), edges=edges, source=np.int32(source),
                       eps=np.float64(c["eps"]), flags=np.array(c["flags"]), variants=np.array(c["variants"], np.int32))
            for v in c["variants"]:
                dump = os.path.join(tmp, f"{name}_{v}.dump")
                run_harness(binpath, c, v, source, dump, scratch=True)
                d = read_dump(dump)
                inc = run_harness(binpath, c, v, source, dump + ".inc", scratch=False)
                snaps = d["snaps"]
                out["W"], out["B"] = np.int64(d["W"]), np.int64(d["B"])
                out[f"v{v}_p"] = np.stack([s["p"] for s in snaps])
                out[f"v{v}_r"] = np.stack([s["r"] for s in snaps])
                out[f"v{v}_iteration_id"] = np.array([s["iteration_id"] for s in snaps], np.int32)
                out[f"v{v}_ref_inc_rows_differ"] = np.int64(inc["harness_inc_rows_differ"])
                if v == c["variants"][0]:  # graph objects do not depend on the variant
                    out["outdeg"] = np.stack([s["outdeg"] for s in snaps])
                    out["in_row_ptr"] = np.stack([s["in_row_ptr"] for s in snaps])
                    out["in_col"] = np.stack([s["in_col"] for s in snaps])
                    out["pow"] = np.stack([s["pow"] for s in snaps])
                    out["n_snap"] = np.int32(len(snaps))
            path = os.path.join(HERE, name + ".npz")
            np.savez_compressed(path, **out)
            print(f"{name}: V={c['V']} M={c['M']} W={int(out['W'])} B={int(out['B'])} snaps={int(out['n_snap'])} "
                  f"source={source} -> {os.path.getsize(path) / 1024:.0f} KiB; unmodified-reference wrong rows: "
                  + ",".join(str(int(out[f'v{v}_ref_inc_rows_differ'])) for v in c["variants"]))


if __name__ == "__main__":
    main()
