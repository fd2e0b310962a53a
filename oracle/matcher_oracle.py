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


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def search_by_bow_kf_f(Dkf, node_kf, good_kf, Df, node_f, ratio=0.7, th_low=100):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...), src/ORBmatcher.cc:408-610 -> (nmatches, matches_f[n_f])."""
    Dkf, p1 = _f(Dkf); Df, p2 = _f(Df)
    nk, pnk = _i(node_kf); nf, pnf = _i(node_f); gk, pgk = _u8(good_kf)
    m = np.empty(Df.shape[0], np.int32)
    L = lib(); L.mo_search_by_bow_kf_f.restype = ctypes.c_int
    n = L.mo_search_by_bow_kf_f(p1, pnk, pgk, Dkf.shape[0], p2, pnf, Df.shape[0], ctypes.c_float(ratio), int(th_low), m.ctypes.data_as(ctypes.c_void_p))
    return n, m


def search_by_bow_kf_kf(D1, node1, good1, D2, node2, good2, ratio=0.9, th_low=100):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...), src/ORBmatcher.cc:950-1090 -> (nmatches, matches12[n1])."""
    D1, p1 = _f(D1); D2, p2 = _f(D2)
    a1, pa1 = _i(node1); a2, pa2 = _i(node2); g1, pg1 = _u8(good1); g2, pg2 = _u8(good2)
    m = np.empty(D1.shape[0], np.int32)
    L = lib(); L.mo_search_by_bow_kf_kf.restype = ctypes.c_int
    n = L.mo_search_by_bow_kf_kf(p1, pa1, pg1, D1.shape[0], p2, pa2, pg2, D2.shape[0], ctypes.c_float(ratio), int(th_low), m.ctypes.data_as(ctypes.c_void_p))
    return n, m


def search_for_triangulation(D1, node1, hasmp1, stereo1, k1xy, D2, node2, hasmp2, stereo2, k2xy, F12, ep, only_stereo=False, coarse=False,
                             th_low=100, unc=1.0, scale0=1.0):
    """ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:1092-1331 (one pinhole camera) -> (nmatches, matches12[n1])."""
    D1, p1 = _f(D1); D2, p2 = _f(D2); k1, pk1 = _f(k1xy); k2, pk2 = _f(k2xy); Fm, pF = _f(np.asarray(F12).reshape(9)); e, pe = _f(ep)
    a1, pa1 = _i(node1); a2, pa2 = _i(node2)
    h1, ph1 = _u8(hasmp1); h2, ph2 = _u8(hasmp2); s1, ps1 = _u8(stereo1); s2, ps2 = _u8(stereo2)
    m = np.empty(D1.shape[0], np.int32)
    L = lib(); L.mo_search_for_triangulation.restype = ctypes.c_int
    n = L.mo_search_for_triangulation(p1, pa1, ph1, ps1, pk1, D1.shape[0], p2, pa2, ph2, ps2, pk2, D2.shape[0], pF, pe, int(only_stereo), int(coarse),
                                      int(th_low), ctypes.c_float(unc), ctypes.c_float(scale0), m.ctypes.data_as(ctypes.c_void_p))
    return n, m


def search_by_projection(Dmp, in_view, proj, projxr, level, viewcos, mp_obs, Df, kxy, occupied, uright, img_w, img_h, th=1.0, scale_factor=1.2,
                         ratio=0.8, th_high=1000):
    """ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th, ...), src/ORBmatcher.cc:42-212 (Nleft == -1) -> (nmatches, assign[n_f])."""
    Dmp, p1 = _f(Dmp); Df, p2 = _f(Df); pr, ppr = _f(proj); px, ppx = _f(projxr); vc, pvc = _f(viewcos); kk, pkk = _f(kxy); ur, pur = _f(uright)
    iv, piv = _u8(in_view); mo_, pmo = _u8(mp_obs); oc, poc = _u8(occupied); lv, plv = _i(level)
    out = np.empty(Df.shape[0], np.int32)
    L = lib(); L.mo_search_by_projection.restype = ctypes.c_int
    n = L.mo_search_by_projection(p1, piv, ppr, ppx, plv, pvc, pmo, Dmp.shape[0], p2, pkk, poc, pur, Df.shape[0], int(img_w), int(img_h),
                                  ctypes.c_float(th), ctypes.c_float(scale_factor), ctypes.c_float(ratio), int(th_high),
                                  out.ctypes.data_as(ctypes.c_void_p))
    return n, out


def distinctive_descriptors(D, offsets):
    """MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:329-403, batched -> best row (relative) per set."""
    D, p = _f(D); off, po = _i(offsets)
    best = np.empty(off.shape[0] - 1, np.int32)
    lib().mo_distinctive_descriptors(p, po, off.shape[0] - 1, best.ctypes.data_as(ctypes.c_void_p))
    return best


def bow_transform(desc, node_desc, child_start, child_index, L, levelsup=4):
    """TemplatedVocabulary::transform per feature (TemplatedVocabulary.h:1218-1260, FORB.cpp:81-101) -> (leaf node, node at L - levelsup)."""
    D, p = _f(np.asarray(desc).reshape(-1, 64))
    nd, pnd = _u8(np.asarray(node_desc).reshape(-1, 32)); cs, pcs = _i(child_start); ci, pci = _i(child_index)
    leaf = np.empty(D.shape[0], np.int32); nid = np.empty(D.shape[0], np.int32)
    lib().mo_bow_transform(p, D.shape[0], pnd, pcs, pci, int(L), int(levelsup), leaf.ctypes.data_as(ctypes.c_void_p), nid.ctypes.data_as(ctypes.c_void_p))
    return leaf, nid


def search_by_projection_frames(Dlast, valid, uv, invzc, octave, mp_obs, Dcur, kxy, occupied, uright, img_w, img_h, th=15.0, scale_factor=1.2,
                                mbf=40.0, forward=False, backward=False, th_high=1000):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono), src/ORBmatcher.cc:1861-2072 -> (nmatches, assign[n_cur])."""
    Dl, p1 = _f(Dlast); Dc, p2 = _f(Dcur); u, pu = _f(uv); iz, piz = _f(invzc); kk, pkk = _f(kxy); ur, pur = _f(uright)
    va, pva = _u8(valid); ob, pob = _u8(mp_obs); oc, poc = _u8(occupied); ot, pot = _i(octave)
    out = np.empty(Dc.shape[0], np.int32)
    L = lib(); L.mo_search_by_projection_frames.restype = ctypes.c_int
    n = L.mo_search_by_projection_frames(p1, pva, pu, piz, pot, pob, Dl.shape[0], p2, pkk, poc, pur, Dc.shape[0], int(img_w), int(img_h),
                                         ctypes.c_float(th), ctypes.c_float(scale_factor), ctypes.c_float(mbf), int(forward), int(backward),
                                         int(th_high), out.ctypes.data_as(ctypes.c_void_p))
    return n, out


def search_by_projection_sim3(Dmp, valid, uv, radius, level, Dkf, kxy, matched, img_w, img_h, th_low=100, ratio_hamming=1.0):
    """ORBmatcher::SearchByProjection(KeyFrame*, Sim3, vpPoints, vpMatched, th, ratioHamming), src/ORBmatcher.cc:612-717 -> (nmatches, assign[n_kf])."""
    Dm, p1 = _f(Dmp); Dk, p2 = _f(Dkf); u, pu = _f(uv); r, pr = _f(radius); kk, pkk = _f(kxy)
    va, pva = _u8(valid); ma, pma = _u8(matched); lv, plv = _i(level)
    out = np.empty(Dk.shape[0], np.int32)
    L = lib(); L.mo_search_by_projection_sim3.restype = ctypes.c_int
    n = L.mo_search_by_projection_sim3(p1, pva, pu, pr, plv, Dm.shape[0], p2, pkk, pma, Dk.shape[0], int(img_w), int(img_h), int(th_low),
                                       ctypes.c_float(ratio_hamming), out.ctypes.data_as(ctypes.c_void_p))
    return n, out


def fuse_search(Dmp, valid, uv, ur, radius, level, Dkf, kxy, uright, img_w, img_h, inv_sigma2_0=1.0):
    """Candidate search of ORBmatcher::Fuse, src/ORBmatcher.cc:1413-1479 -> (best_idx[n_mp], best_dist[n_mp])."""
    Dm, p1 = _f(Dmp); Dk, p2 = _f(Dkf); u, pu = _f(uv); urr, pur = _f(ur); r, pr = _f(radius); kk, pkk = _f(kxy); rr, prr = _f(uright)
    va, pva = _u8(valid); lv, plv = _i(level)
    bi = np.empty(Dm.shape[0], np.int32); bd = np.empty(Dm.shape[0], np.int32)
    lib().mo_fuse_search(p1, pva, pu, pur, pr, plv, Dm.shape[0], p2, pkk, prr, Dk.shape[0], int(img_w), int(img_h), ctypes.c_float(inv_sigma2_0),
                         bi.ctypes.data_as(ctypes.c_void_p), bd.ctypes.data_as(ctypes.c_void_p))
    return bi, bd


def search_by_sim3(Dmp1, valid1, uv1, radius1, level1, D2, k2xy, Dmp2, valid2, uv2, radius2, level2, D1, k1xy, img_w, img_h, th_high=1000):
    """ORBmatcher::SearchBySim3, src/ORBmatcher.cc:1642-1859 -> (nFound, matches12[n1])."""
    a1, pa1 = _f(Dmp1); a2, pa2 = _f(Dmp2); d1, pd1 = _f(D1); d2, pd2 = _f(D2)
    u1, pu1 = _f(uv1); u2, pu2 = _f(uv2); r1, pr1 = _f(radius1); r2, pr2 = _f(radius2); k1, pk1 = _f(k1xy); k2, pk2 = _f(k2xy)
    v1, pv1 = _u8(valid1); v2, pv2 = _u8(valid2); l1, pl1 = _i(level1); l2, pl2 = _i(level2)
    out = np.empty(a1.shape[0], np.int32)
    L = lib(); L.mo_search_by_sim3.restype = ctypes.c_int
    n = L.mo_search_by_sim3(pa1, pv1, pu1, pr1, pl1, a1.shape[0], pd2, pk2, pa2, pv2, pu2, pr2, pl2, a2.shape[0], pd1, pk1, int(img_w), int(img_h),
                            int(th_high), out.ctypes.data_as(ctypes.c_void_p))
    return n, out


def search_by_projection_reloc(Dmp, valid, uv, radius, level, Dcur, kxy, occupied, img_w, img_h, orb_dist=100):
    """ORBmatcher::SearchByProjection(CurrentFrame, KeyFrame*, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:2074-2190 -> (nmatches, assign)."""
    Dm, p1 = _f(Dmp); Dc, p2 = _f(Dcur); u, pu = _f(uv); r, pr = _f(radius); kk, pkk = _f(kxy)
    va, pva = _u8(valid); oc, poc = _u8(occupied); lv, plv = _i(level)
    out = np.empty(Dc.shape[0], np.int32)
    L = lib(); L.mo_search_by_projection_reloc.restype = ctypes.c_int
    n = L.mo_search_by_projection_reloc(p1, pva, pu, pr, plv, Dm.shape[0], p2, pkk, poc, Dc.shape[0], int(img_w), int(img_h), int(orb_dist),
                                        out.ctypes.data_as(ctypes.c_void_p))
    return n, out
