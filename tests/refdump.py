"""Reader for the binary dump written by oracle/ref_harness.cpp (test infrastructure)."""
from __future__ import annotations
import struct
import numpy as np


def read_dump(path: str) -> dict:
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == b"DPPRDMP1", "bad dump magic"
    V, directed, W, B, has_pow, nsnap = struct.unpack_from("<iiqqii", buf, 8)
    off = 8 + struct.calcsize("<iiqqii")
    snaps = []
    for _ in range(nsnap):
        batch_index, iteration_id, E, ppr_us, inc_rows_differ = struct.unpack_from("<iiqdi", buf, off)
        off += struct.calcsize("<iiqdi")

        def take(dtype, n):
            nonlocal off
            a = np.frombuffer(buf, dtype=dtype, count=n, offset=off).copy()
            off += a.nbytes
            return a

        s = dict(batch_index=batch_index, iteration_id=iteration_id, E=E, ppr_us=ppr_us,
                 inc_rows_differ=inc_rows_differ)
        s["p"] = take(np.float64, V)
        s["r"] = take(np.float64, V)
        s["outdeg"] = take(np.int32, V)
        s["in_row_ptr"] = take(np.int32, V + 1)
        s["in_col"] = take(np.int32, E)
        if has_pow:
            s["pow"] = take(np.float64, V)
        snaps.append(s)
    assert off == len(buf), "trailing bytes in dump"
    return dict(V=V, directed=bool(directed), W=W, B=B, has_pow=bool(has_pow), snaps=snaps)
