"""The C restatements of the node-gated / windowed matchers (oracle/matcher_oracle.c) against an independent, deliberately
naive Python transcription of the same reference functions (dict-of-lists FeatureVector like std::map, per-pair distance
calls), on small random problems.  Guards the merge join / CSR / grid code of the C oracle; runs on CPU."""
import math

import numpy as np
import pytest

from oracle import matcher_oracle as mo


def unit(rng, n):
    v = rng.randn(n, 64).astype(np.float32)
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def related(rng, A, idx, noise):
    B = A[idx] + noise * rng.randn(len(idx), 64).astype(np.float32)
    return (B / np.linalg.norm(B, axis=1, keepdims=True)).astype(np.float32)


def featvec(node):
    fv = {}
    for i, nd in enumerate(node):
        if nd >= 0:
            fv.setdefault(int(nd), []).append(i)
    return fv


def problem(seed, n1=160, n2=170, nodes=12):
    rng = np.random.RandomState(seed)
    A = unit(rng, n1)
    src = rng.randint(0, n1, n2)
    B = related(rng, A, src, 0.02)
    fresh = rng.rand(n2) < 0.3
    B[fresh] = unit(rng, int(fresh.sum()))
    nodeA = rng.randint(-1, nodes, n1).astype(np.int32)
    nodeB = np.where(rng.rand(n2) < 0.8, nodeA[src], rng.randint(-1, nodes, n2)).astype(np.int32)
    problem.src = src
    return rng, A, B, nodeA, nodeB


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("ratio", [0.7, 0.9])
def test_search_by_bow_kf_f(seed, ratio):
    rng, A, B, nodeA, nodeB = problem(seed)
    good = (rng.rand(len(A)) < 0.8).astype(np.uint8)
    n, m = mo.search_by_bow_kf_f(A, nodeA, good, B, nodeB, ratio=ratio, th_low=100)
    # src/ORBmatcher.cc:408-610 transcribed naively
    fa, fb = featvec(nodeA), featvec(nodeB)
    want = np.full(len(B), -1, np.int32); cnt = 0
    for node in sorted(set(fa) & set(fb)):
        for i in fa[node]:
            if not good[i]:
                continue
            b1, b2, bi = 256, 256, -1
            for j in fb[node]:
                if want[j] >= 0:
                    continue
                d = mo.descriptor_distance(A[i], B[j])
                if d < b1:
                    b2, b1, bi = b1, d, j
                elif d < b2:
                    b2 = d
            if b1 <= 100 and np.float32(b1) < np.float32(ratio) * np.float32(b2):
                want[bi] = i; cnt += 1
    assert n == cnt and np.array_equal(m, want) and cnt > 10


@pytest.mark.parametrize("seed", [3, 4])
def test_search_by_bow_kf_kf(seed):
    rng, A, B, nodeA, nodeB = problem(seed)
    g1 = (rng.rand(len(A)) < 0.8).astype(np.uint8); g2 = (rng.rand(len(B)) < 0.8).astype(np.uint8)
    n, m = mo.search_by_bow_kf_kf(A, nodeA, g1, B, nodeB, g2, ratio=0.9, th_low=100)
    fa, fb = featvec(nodeA), featvec(nodeB)
    want = np.full(len(A), -1, np.int32); used = np.zeros(len(B), bool); cnt = 0
    for node in sorted(set(fa) & set(fb)):
        for i in fa[node]:
            if not g1[i]:
                continue
            b1, b2, bi = 256, 256, -1
            for j in fb[node]:
                if used[j] or not g2[j]:
                    continue
                d = mo.descriptor_distance(A[i], B[j])
                if d < b1:
                    b2, b1, bi = b1, d, j
                elif d < b2:
                    b2 = d
            if b1 < 100 and np.float32(b1) < np.float32(0.9) * np.float32(b2):
                want[i] = bi; used[bi] = True; cnt += 1
    assert n == cnt and np.array_equal(m, want) and cnt > 10


@pytest.mark.parametrize("coarse", [False, True])
def test_search_for_triangulation(coarse):
    rng, A, B, nodeA, nodeB = problem(5)
    h1 = (rng.rand(len(A)) < 0.3).astype(np.uint8); h2 = (rng.rand(len(B)) < 0.3).astype(np.uint8)
    s1 = (rng.rand(len(A)) < 0.5).astype(np.uint8); s2 = (rng.rand(len(B)) < 0.5).astype(np.uint8)
    k1 = (rng.rand(len(A), 2) * [640, 480]).astype(np.float32); k2 = (rng.rand(len(B), 2) * [640, 480]).astype(np.float32)
    # pure translation along x: F = [t]x with t = (1, 0, 0) -> epipolar lines are horizontal (y2 == y1)
    F = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32)
    k2[:, 1] = np.where(rng.rand(len(B)) < 0.7, k1[problem.src, 1], k2[:, 1])   # most true correspondences lie on their epipolar line
    ep = np.array([320.0, 240.0], np.float32)
    n, m = mo.search_for_triangulation(A, nodeA, h1, s1, k1, B, nodeB, h2, s2, k2, F, ep, only_stereo=False, coarse=coarse)
    fa, fb = featvec(nodeA), featvec(nodeB)
    want = np.full(len(A), -1, np.int32); cnt = 0
    f32 = np.float32
    for node in sorted(set(fa) & set(fb)):
        for i in fa[node]:
            if h1[i]:
                continue
            best, bi = 100, -1
            for j in fb[node]:
                if h2[j]:
                    continue
                d = mo.descriptor_distance(A[i], B[j])
                if d > 100 or d > best:
                    continue
                if not s1[i] and not s2[j]:
                    ex, ey = f32(ep[0] - k2[j, 0]), f32(ep[1] - k2[j, 1])
                    if f32(f32(ex * ex) + f32(ey * ey)) < 100:
                        continue
                ok = coarse
                if not ok:
                    x1, y1, x2, y2 = k1[i, 0], k1[i, 1], k2[j, 0], k2[j, 1]
                    la = f32(f32(f32(x1 * F[0, 0]) + f32(y1 * F[1, 0])) + F[2, 0])
                    lb = f32(f32(f32(x1 * F[0, 1]) + f32(y1 * F[1, 1])) + F[2, 1])
                    lc = f32(f32(f32(x1 * F[0, 2]) + f32(y1 * F[1, 2])) + F[2, 2])
                    num = f32(f32(f32(la * x2) + f32(lb * y2)) + lc)
                    den = f32(f32(la * la) + f32(lb * lb))
                    ok = den != 0 and float(f32(f32(num * num) / den)) < 3.84 * 1.0
                if ok:
                    bi, best = j, d
            if bi >= 0:
                want[i] = bi; cnt += 1
    assert n == cnt and np.array_equal(m, want) and cnt > 5


def test_search_by_projection():
    rng = np.random.RandomState(11)
    nF, nM, W, H = 400, 150, 640, 480
    kxy = np.stack([rng.randint(0, W, nF), rng.randint(0, H, nF)], 1).astype(np.float32)
    Df = unit(rng, nF)
    src = rng.randint(0, nF, nM)
    Dmp = related(rng, Df, src, 0.03)
    proj = (kxy[src] + rng.randn(nM, 2) * 1.5).astype(np.float32)
    in_view = (rng.rand(nM) < 0.9).astype(np.uint8)
    level = rng.choice([0, 0, 0, 1, 2], nM).astype(np.int32)
    viewcos = rng.choice([0.9, 0.9995], nM).astype(np.float32)
    mp_obs = (rng.rand(nM) < 0.9).astype(np.uint8)
    occupied = (rng.rand(nF) < 0.1).astype(np.uint8)
    uright = np.where(rng.rand(nF) < 0.5, kxy[:, 0] - 20.0, -1.0).astype(np.float32)
    projxr = (proj[:, 0] - 20.0 + rng.randn(nM) * 3).astype(np.float32)
    n, a = mo.search_by_projection(Dmp, in_view, proj, projxr, level, viewcos, mp_obs, Df, kxy, occupied, uright, W, H, th=3.0, scale_factor=1.2,
                                   ratio=0.8, th_high=1000)
    # naive transcription of src/ORBmatcher.cc:42-141 with a brute-force GetFeaturesInArea (cell order = ix, iy, insertion)
    f32 = np.float32
    wInv, hInv = f32(64) / f32(W), f32(48) / f32(H)
    cells = {}
    for i in range(nF):
        px, py = int(np.round(f32(kxy[i, 0] * wInv))), int(np.round(f32(kxy[i, 1] * hInv)))
        px = int(math.floor(float(f32(kxy[i, 0] * wInv)) + 0.5)) if kxy[i, 0] >= 0 else px   # C round(): half away from zero
        py = int(math.floor(float(f32(kxy[i, 1] * hInv)) + 0.5)) if kxy[i, 1] >= 0 else py
        if 0 <= px < 64 and 0 <= py < 48:
            cells.setdefault((px, py), []).append(i)
    occ = occupied.copy(); want = np.full(nF, -1, np.int32); cnt = 0
    for m in range(nM):
        if not in_view[m]:
            continue
        lvl = int(level[m])
        r = f32(2.5) if viewcos[m] > f32(0.998) else f32(4.0)
        r = f32(r * f32(3.0))
        sf = f32(1.0)
        for _ in range(lvl):
            sf = f32(sf * f32(1.2))
        rr = f32(r * sf)
        if lvl - 1 > 0:
            continue                                   # octave 0 < minLevel: no candidates
        x, y = proj[m]
        c0 = max(0, int(math.floor(float(f32(f32(x - rr) * wInv))))); c1 = min(63, int(math.ceil(float(f32(f32(x + rr) * wInv)))))
        r0 = max(0, int(math.floor(float(f32(f32(y - rr) * hInv))))); r1 = min(47, int(math.ceil(float(f32(f32(y + rr) * hInv)))))
        if c0 >= 64 or c1 < 0 or r0 >= 48 or r1 < 0:
            continue
        cand = [j for ix in range(c0, c1 + 1) for iy in range(r0, r1 + 1) for j in cells.get((ix, iy), [])
                if abs(f32(kxy[j, 0] - x)) < rr and abs(f32(kxy[j, 1] - y)) < rr]
        if not cand:
            continue
        b1, l1, b2, l2, bi = 256, -1, 256, -1, -1
        for j in cand:
            if occ[j]:
                continue
            if uright[j] > 0 and abs(f32(projxr[m] - uright[j])) > rr:
                continue
            d = mo.descriptor_distance(Dmp[m], Df[j])
            if d < b1:
                b2, b1, l2, l1, bi = b1, d, l1, 0, j
            elif d < b2:
                l2, b2 = 0, d
        if b1 <= 1000:
            if l1 == l2 and f32(b1) > f32(0.8) * f32(b2):
                continue
            want[bi] = m; occ[bi] = mp_obs[m]; cnt += 1
    assert n == cnt and np.array_equal(a, want) and cnt > 30


def test_distinctive_descriptors():
    rng = np.random.RandomState(12)
    sizes = [1, 2, 3, 7, 8, 30]
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    base = unit(rng, len(sizes))
    D = np.concatenate([related(rng, base, [s] * n, 0.05) for s, n in enumerate(sizes)])
    best = mo.distinctive_descriptors(D, off)
    for s, n in enumerate(sizes):
        d = mo.distance_matrix(D[off[s]:off[s + 1]], D[off[s]:off[s + 1]])
        np.fill_diagonal(d, 0)
        med = [sorted(d[i])[int(0.5 * (n - 1))] for i in range(n)]
        assert best[s] == int(np.argmin(med))           # first minimum wins (:388-393)


def _naive_area(kxy, cells, W, H, x, y, rr, min_level, max_level):
    """Frame::GetFeaturesInArea (src/Frame.cc:850-916) for octave-0 keypoints, brute force over the cell dictionary."""
    f32 = np.float32
    if (min_level > 0 or max_level >= 0) and 0 < min_level:
        return []
    wInv, hInv = f32(64) / f32(W), f32(48) / f32(H)
    c0 = max(0, int(math.floor(float(f32(f32(x - rr) * wInv))))); c1 = min(63, int(math.ceil(float(f32(f32(x + rr) * wInv)))))
    r0 = max(0, int(math.floor(float(f32(f32(y - rr) * hInv))))); r1 = min(47, int(math.ceil(float(f32(f32(y + rr) * hInv)))))
    if c0 >= 64 or c1 < 0 or r0 >= 48 or r1 < 0:
        return []
    return [j for ix in range(c0, c1 + 1) for iy in range(r0, r1 + 1) for j in cells.get((ix, iy), [])
            if abs(f32(kxy[j, 0] - x)) < rr and abs(f32(kxy[j, 1] - y)) < rr]


@pytest.mark.parametrize("forward,backward", [(False, False), (True, False), (False, True)])
def test_search_by_projection_frames(forward, backward):
    rng = np.random.RandomState(17)
    nL, nC, W, H = 300, 420, 640, 480
    kxy = np.stack([rng.randint(0, W, nC), rng.randint(0, H, nC)], 1).astype(np.float32)
    Dc = unit(rng, nC)
    src = rng.randint(0, nC, nL)
    Dl = related(rng, Dc, src, 0.03)
    Dl[rng.rand(nL) < 0.2] = unit(rng, 1)[0]                      # map points that match nothing well (distance >= 256 everywhere)
    uv = (kxy[src] + rng.randn(nL, 2) * 3).astype(np.float32)
    valid = (rng.rand(nL) < 0.85).astype(np.uint8)
    invzc = (1.0 / (0.5 + 4 * rng.rand(nL))).astype(np.float32)
    octave = rng.choice([0, 0, 0, 1, 2], nL).astype(np.int32)
    obs = (rng.rand(nL) < 0.9).astype(np.uint8)
    occupied = (rng.rand(nC) < 0.1).astype(np.uint8)
    uright = np.where(rng.rand(nC) < 0.5, kxy[:, 0] - 20.0, -1.0).astype(np.float32)
    n, a = mo.search_by_projection_frames(Dl, valid, uv, invzc, octave, obs, Dc, kxy, occupied, uright, W, H, th=15.0, scale_factor=1.2, mbf=40.0,
                                          forward=forward, backward=backward, th_high=1000)
    f32 = np.float32
    wInv, hInv = f32(64) / f32(W), f32(48) / f32(H)
    cells = {}
    for i in range(nC):
        px, py = int(math.floor(float(f32(kxy[i, 0] * wInv)) + 0.5)), int(math.floor(float(f32(kxy[i, 1] * hInv)) + 0.5))
        if 0 <= px < 64 and 0 <= py < 48:
            cells.setdefault((px, py), []).append(i)
    occ = occupied.copy(); want = np.full(nC, -1, np.int32); cnt = 0
    for i in range(nL):
        if not valid[i]:
            continue
        o = int(octave[i])
        sf = f32(1.0)
        for _ in range(o):
            sf = f32(sf * f32(1.2))
        radius = f32(f32(15.0) * sf)
        lv = (o, -1) if forward else ((0, o) if backward else (o - 1, o + 1))
        cand = _naive_area(kxy, cells, W, H, uv[i, 0], uv[i, 1], radius, *lv)
        best, bi = 256, -1
        for j in cand:
            if occ[j]:
                continue
            if uright[j] > 0 and abs(f32(f32(uv[i, 0] - f32(f32(40.0) * invzc[i])) - uright[j])) > radius:
                continue
            d = mo.descriptor_distance(Dl[i], Dc[j])
            if d < best:
                best, bi = d, j
        if best <= 1000 and bi >= 0:                                # (the reference omits `bi >= 0`: out-of-bounds write, see the oracle)
            want[bi] = i; occ[bi] = obs[i]; cnt += 1
    assert n == cnt and np.array_equal(a, want) and cnt > 30


def _window_problem(seed):
    rng = np.random.RandomState(seed)
    nM, nK, W, H = 350, 450, 640, 480
    kxy = np.stack([rng.randint(0, W, nK), rng.randint(0, H, nK)], 1).astype(np.float32)
    Dk = unit(rng, nK)
    src = rng.randint(0, nK, nM)
    Dm = related(rng, Dk, src, 0.04)
    uv = (kxy[src] + rng.randn(nM, 2) * 1.5).astype(np.float32)
    valid = (rng.rand(nM) < 0.85).astype(np.uint8)
    level = rng.choice([0, 0, 1, 2], nM).astype(np.int32)
    radius = (np.float32(5.0) * np.float32(1.2) ** level).astype(np.float32)
    cells = {}
    f32 = np.float32
    wInv, hInv = f32(64) / f32(W), f32(48) / f32(H)
    for i in range(nK):
        px, py = int(math.floor(float(f32(kxy[i, 0] * wInv)) + 0.5)), int(math.floor(float(f32(kxy[i, 1] * hInv)) + 0.5))
        if 0 <= px < 64 and 0 <= py < 48:
            cells.setdefault((px, py), []).append(i)
    return rng, Dm, Dk, kxy, uv, valid, level, radius, cells, W, H


def test_search_by_projection_sim3():
    rng, Dm, Dk, kxy, uv, valid, level, radius, cells, W, H = _window_problem(23)
    matched = (rng.rand(len(Dk)) < 0.1).astype(np.uint8)
    n, a = mo.search_by_projection_sim3(Dm, valid, uv, radius, level, Dk, kxy, matched, W, H, th_low=100, ratio_hamming=0.9)
    occ = matched.copy(); want = np.full(len(Dk), -1, np.int32); cnt = 0
    for m in range(len(Dm)):
        if not valid[m]:
            continue
        best, bi = 256, -1
        for j in _naive_area(kxy, cells, W, H, uv[m, 0], uv[m, 1], radius[m], -1, -1):
            if occ[j] or 0 < level[m] - 1 or 0 > level[m]:
                continue
            d = mo.descriptor_distance(Dm[m], Dk[j])
            if d < best:
                best, bi = d, j
        if np.float32(best) <= np.float32(100) * np.float32(0.9):
            want[bi] = m; occ[bi] = 1; cnt += 1
    assert n == cnt and np.array_equal(a, want) and cnt > 30


def test_fuse_search():
    rng, Dm, Dk, kxy, uv, valid, level, radius, cells, W, H = _window_problem(29)
    uright = np.where(rng.rand(len(Dk)) < 0.5, kxy[:, 0] - 20.0, -1.0).astype(np.float32)
    ur = (uv[:, 0] - 20.0 + rng.randn(len(Dm))).astype(np.float32)
    bi, bd = mo.fuse_search(Dm, valid, uv, ur, radius, level, Dk, kxy, uright, W, H, inv_sigma2_0=1.0)
    f32 = np.float32
    wbi = np.full(len(Dm), -1, np.int32); wbd = np.full(len(Dm), 256, np.int32)
    for m in range(len(Dm)):
        if not valid[m]:
            continue
        for j in _naive_area(kxy, cells, W, H, uv[m, 0], uv[m, 1], radius[m], -1, -1):
            if 0 < level[m] - 1 or 0 > level[m]:
                continue
            ex, ey = f32(uv[m, 0] - kxy[j, 0]), f32(uv[m, 1] - kxy[j, 1])
            if uright[j] >= 0:
                er = f32(ur[m] - uright[j])
                if float(f32(f32(f32(ex * ex) + f32(ey * ey)) + f32(er * er)) * f32(1.0)) > 7.8:
                    continue
            elif float(f32(f32(ex * ex) + f32(ey * ey)) * f32(1.0)) > 5.99:
                continue
            d = mo.descriptor_distance(Dm[m], Dk[j])
            if d < wbd[m]:
                wbd[m], wbi[m] = d, j
    assert np.array_equal(bi, wbi) and np.array_equal(bd, wbd) and (bd <= 100).sum() > 30


def test_search_by_sim3():
    rng, Dm, Dk, kxy, uv, valid, level, radius, cells, W, H = _window_problem(31)
    # keyframe 2 = (Dk, kxy); keyframe 1 = the "map points" with their own positions = the projections moved back by (4, -2)
    k1 = (uv - np.array([4, -2], np.float32)).astype(np.float32)
    k1 = np.clip(k1, 0, [W - 1, H - 1]).astype(np.float32)
    n1, n2 = len(Dm), len(Dk)
    uv2 = (kxy - np.array([4, -2], np.float32) + rng.randn(n2, 2)).astype(np.float32)        # KF2 points projected into KF1
    valid2 = (rng.rand(n2) < 0.85).astype(np.uint8)
    level2 = rng.choice([0, 0, 1, 2], n2).astype(np.int32)
    radius2 = (np.float32(7.5) * np.float32(1.2) ** level2).astype(np.float32)
    radius1 = (np.float32(7.5) * np.float32(1.2) ** level).astype(np.float32)
    n, m12 = mo.search_by_sim3(Dm, valid, uv, radius1, level, Dk, kxy, Dk, valid2, uv2, radius2, level2, Dm, k1, W, H, th_high=1000)
    f32 = np.float32
    wInv, hInv = f32(64) / f32(W), f32(48) / f32(H)
    cells1 = {}
    for i in range(n1):
        px, py = int(math.floor(float(f32(k1[i, 0] * wInv)) + 0.5)), int(math.floor(float(f32(k1[i, 1] * hInv)) + 0.5))
        if 0 <= px < 64 and 0 <= py < 48:
            cells1.setdefault((px, py), []).append(i)

    def one_way(Dq, ok, uvq, rad, lvl, Dt, kt, ct):
        out = np.full(len(Dq), -1, np.int64)
        for i in range(len(Dq)):
            if not ok[i]:
                continue
            best, bi = 2 ** 31 - 1, -1
            for j in _naive_area(kt, ct, W, H, uvq[i, 0], uvq[i, 1], rad[i], -1, -1):
                if 0 < lvl[i] - 1 or 0 > lvl[i]:
                    continue
                d = mo.descriptor_distance(Dq[i], Dt[j])
                if d < best:
                    best, bi = d, j
            if best <= 1000:
                out[i] = bi
        return out
    a = one_way(Dm, valid, uv, radius1, level, Dk, kxy, cells)
    b = one_way(Dk, valid2, uv2, radius2, level2, Dm, k1, cells1)
    want = np.array([a[i] if a[i] >= 0 and b[a[i]] == i else -1 for i in range(n1)], np.int32)
    assert n == int((want >= 0).sum()) and np.array_equal(m12, want) and n > 20


def test_search_by_projection_reloc():
    rng, Dm, Dk, kxy, uv, valid, level, radius, cells, W, H = _window_problem(37)
    occupied = (rng.rand(len(Dk)) < 0.15).astype(np.uint8)
    n, a = mo.search_by_projection_reloc(Dm, valid, uv, radius, level, Dk, kxy, occupied, W, H, orb_dist=64)
    occ = occupied.copy(); want = np.full(len(Dk), -1, np.int32); cnt = 0
    for m in range(len(Dm)):
        if not valid[m]:
            continue
        best, bi = 256, -1
        for j in _naive_area(kxy, cells, W, H, uv[m, 0], uv[m, 1], radius[m], level[m] - 1, level[m] + 1):
            if occ[j]:
                continue
            d = mo.descriptor_distance(Dm[m], Dk[j])
            if d < best:
                best, bi = d, j
        if best <= 64:
            want[bi] = m; occ[bi] = 1; cnt += 1
    assert n == cnt and np.array_equal(a, want) and cnt > 30
