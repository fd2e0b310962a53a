"""ctypes wrapper around oracle/matcher_oracle.c -- TEST INFRASTRUCTURE ONLY (see the C header)."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libmatcher_oracle.so"
_lib = None


def build(force=False):
    src = _HERE / "matcher_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        _SO.parent.mkdir(exist_ok=True)
        subprocess.run(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off", str(src), "-o", str(_SO), "-lm"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
        _lib.mo_descriptor_distance.restype = ctypes.c_int
        _lib.mo_search_for_initialization.restype = ctypes.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def descriptor_distance(a, b):
    a, pa = _f(a); b, pb = _f(b)
    return lib().mo_descriptor_distance(pa, pb)


def distance_matrix(A, B):
    A, pa = _f(A); B, pb = _f(B)
    out = np.empty((A.shape[0], B.shape[0]), np.int32)
    lib().mo_distance_matrix(pa, A.shape[0], pb, B.shape[0], out.ctypes.data_as(ctypes.c_void_p))
    return out


def bruteforce(A, B, group_a=None, group_b=None, init=2 ** 31 - 1):
    A, pa = _f(A); B, pb = _f(B)
    n1, n2 = A.shape[0], B.shape[0]
    ga = gb = None
    pga = pgb = None
    if group_a is not None:
        ga, pga = _i(group_a); gb, pgb = _i(group_b)
    bi = np.empty(n1, np.int32); bd = np.empty(n1, np.int32); sd = np.empty(n1, np.int32)
    ri = np.empty(n2, np.int32); rd = np.empty(n2, np.int32)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    lib().mo_bruteforce(pa, n1, pb, n2, pga, pgb, ctypes.c_int(int(init)), p(bi), p(bd), p(sd), p(ri), p(rd))
    return bi, bd, sd, ri, rd


def search_for_initialization(D1, k1xy, D2, k2xy, img_w, img_h, prev, window=100, ratio=0.9, th_low=100):
    D1, p1 = _f(D1); D2, p2 = _f(D2)
    k1, pk1 = _f(k1xy); k2, pk2 = _f(k2xy)
    prev = np.array(prev, dtype=np.float32, copy=True)
    m = np.empty(D1.shape[0], np.int32)
    n = lib().mo_search_for_initialization(p1, pk1, D1.shape[0], p2, pk2, D2.shape[0], int(img_w), int(img_h),
                                           prev.ctypes.data_as(ctypes.c_void_p), int(window), ctypes.c_float(ratio), int(th_low),
                                           m.ctypes.data_as(ctypes.c_void_p))
    return n, m, prev
