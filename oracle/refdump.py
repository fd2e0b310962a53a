"""Reader for the record stream written by oracle/_ref/ref_xfeat dump -- TEST INFRASTRUCTURE ONLY."""
import struct

import numpy as np

_DT = {0: np.dtype("<f4"), 1: np.dtype("<i8"), 2: np.dtype("u1"), 3: np.dtype("<i4")}


def read_dump(path):
    raw = open(path, "rb").read()
    out, pos = {}, 0
    while pos < len(raw):
        (nlen,) = struct.unpack_from("<I", raw, pos); pos += 4
        name = raw[pos:pos + nlen].decode(); pos += nlen
        code, nd = struct.unpack_from("<II", raw, pos); pos += 8
        dims = struct.unpack_from("<%dq" % nd, raw, pos); pos += 8 * nd
        dt = _DT[code]
        n = int(np.prod(dims)) if nd else 1
        out[name] = np.frombuffer(raw, dtype=dt, count=n, offset=pos).reshape(dims).copy()
        pos += n * dt.itemsize
    return out
