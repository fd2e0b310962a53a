#!/usr/bin/env python
"""Convert the reference's weight archive into the flat blob libxfeat_b200.so loads.

Reads /root/reference/weights/xfeat.pt (the TorchScript-style archive the reference loads with
torch::serialize::InputArchive, src/XFextractor.cc:133-137) and writes
xfeatslam_b200/weights/xfeat_b200.bin.  Only the tensors the hot path reads are kept: the conv
weights/biases of SURVEY.md section 8a row W.  The BatchNorm running statistics are dropped on
purpose -- the reference runs BN in training mode (never calls eval(), src/XFextractor.cc:133-144),
so they never influence an output; `fine_matcher` is never called (src/XFeat.cc:135-173).

Blob layout (little endian):
  header : magic 'XFBW' | u32 version=1 | u32 n_tensors | u32 reserved
  table  : n_tensors x { char name[48]; u32 ndim; u32 dims[4]; u64 offset; u64 nbytes }
  data   : float32 tensors in the reference's OIHW order, each 64-byte aligned
"""
import argparse
import struct
import sys
from pathlib import Path

import numpy as np

MAGIC = b"XFBW"
VERSION = 1
ENTRY = struct.Struct("<48sI4IQQ")


def collect(src):
    import torch

    mod = torch.jit.load(src, map_location="cpu")
    keep = []
    for name, p in mod.named_parameters():
        if name.startswith("fine_matcher"):
            continue
        keep.append((name, p.detach().to(torch.float32).contiguous().numpy()))
    return keep


def write_blob(tensors, dst):
    n = len(tensors)
    table_bytes = 16 + n * ENTRY.size
    off = (table_bytes + 63) // 64 * 64
    entries, chunks = [], []
    for name, arr in tensors:
        assert arr.dtype == np.float32 and arr.ndim <= 4
        dims = list(arr.shape) + [1] * (4 - arr.ndim)
        nbytes = arr.nbytes
        entries.append(ENTRY.pack(name.encode(), arr.ndim, *dims, off, nbytes))
        chunks.append((off, arr.tobytes()))
        off = (off + nbytes + 63) // 64 * 64
    out = bytearray(off)
    out[0:16] = MAGIC + struct.pack("<III", VERSION, n, 0)
    pos = 16
    for e in entries:
        out[pos:pos + ENTRY.size] = e
        pos += ENTRY.size
    for o, b in chunks:
        out[o:o + len(b)] = b
    Path(dst).parent.mkdir(parents=True, exist_ok=True)
    Path(dst).write_bytes(bytes(out))
    return off


def read_blob(path):
    """Returns {name: np.ndarray(float32)} -- shared by the oracle and the ctypes binding."""
    raw = Path(path).read_bytes()
    if raw[:4] != MAGIC:
        raise ValueError("not an XFBW blob: %s" % path)
    version, n, _ = struct.unpack_from("<III", raw, 4)
    if version != VERSION:
        raise ValueError("unsupported blob version %d" % version)
    out = {}
    pos = 16
    for _ in range(n):
        name, ndim, d0, d1, d2, d3, off, nbytes = ENTRY.unpack_from(raw, pos)
        pos += ENTRY.size
        shape = (d0, d1, d2, d3)[:ndim]
        arr = np.frombuffer(raw, dtype="<f4", count=nbytes // 4, offset=off).reshape(shape).copy()
        out[name.rstrip(b"\0").decode()] = arr
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference/weights/xfeat.pt")
    ap.add_argument("--dst", default=str(Path(__file__).resolve().parents[1] / "xfeatslam_b200" / "weights" / "xfeat_b200.bin"))
    a = ap.parse_args()
    tensors = collect(a.src)
    size = write_blob(tensors, a.dst)
    back = read_blob(a.dst)
    for name, arr in tensors:
        assert np.array_equal(back[name], arr), name
    nparam = sum(t.size for _, t in tensors)
    print("wrote %s: %d tensors, %d params, %d bytes" % (a.dst, len(tensors), nparam, size))
    return 0


if __name__ == "__main__":
    sys.exit(main())
