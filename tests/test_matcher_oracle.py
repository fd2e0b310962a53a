"""oracle/matcher_oracle.c: pinned against OpenCV's cv::norm (python-opencv in this image) and
checked for the reference's scan semantics (src/ORBmatcher.cc:476-486, :833-948)."""
import numpy as np
import pytest

from oracle import matcher_oracle as mo


def unit(rng, n):
    a = rng.randn(n, 64).astype(np.float32)
    return a / np.linalg.norm(a, axis=1, keepdims=True)


def test_distance_against_cv2_norm():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(1)
    A = unit(rng, 300)
    B = (0.8 * A[rng.permutation(300)] + 0.6 * unit(rng, 300)).astype(np.float32)
    B /= np.linalg.norm(B, axis=1, keepdims=True)
    M = mo.distance_matrix(A, B)
    flips = 0
    for i in range(0, 300, 3):
        for j in range(0, 300, 5):
            d = int(np.float32(cv2.norm(A[i:i + 1], B[j:j + 1], cv2.NORM_L2SQR)) * np.float32(512))
            assert abs(d - M[i, j]) <= 1
            flips += d != M[i, j]
    # accumulation order is build dependent: tolerate the documented ~1e-4 flip rate
    assert flips <= 3


def test_distance_known_answers():
    z = np.zeros(64, np.float32)
    e0 = z.copy(); e0[0] = 1
    e1 = z.copy(); e1[1] = 1
    assert mo.descriptor_distance(e0, e0) == 0
    assert mo.descriptor_distance(z, e0) == 512          # phantom row vs unit row
    assert mo.descriptor_distance(z, z) == 0
    assert mo.descriptor_distance(e0, e1) == 1024
    assert mo.descriptor_distance(e0, -e0) == 2048


def test_bruteforce_scan_rule():
    rng = np.random.RandomState(2)
    A, B = unit(rng, 50), unit(rng, 70)
    B[10] = B[3]                                          # duplicate column -> tie, lowest index wins
    A[0] = B[3]
    M = mo.distance_matrix(A, B).astype(np.int64)
    for init in (2 ** 31 - 1, 256):
        bi, bd, sd, ri, rd = mo.bruteforce(A, B, init=init)
        for i in range(50):
            b1, b2, idx = init, init, -1
            for j in range(70):
                d = M[i, j]
                if d < b1:
                    b2, b1, idx = b1, d, j
                elif d < b2:
                    b2 = d
            assert (bi[i], bd[i], sd[i]) == (idx, b1, b2)
        assert bi[0] == 3 and bd[0] == 0 and sd[0] == 0
        for j in range(70):
            col = M[:, j]
            k = int(np.argmin(col))
            if col[k] < init:
                assert ri[j] == k and rd[j] == col[k]
            else:
                assert ri[j] == -1


def test_search_for_initialization_basic():
    rng = np.random.RandomState(3)
    D1 = unit(rng, 200)
    perm = rng.permutation(200)
    D2 = D1[perm] + 0.02 * rng.randn(200, 64).astype(np.float32)
    D2 = (D2 / np.linalg.norm(D2, axis=1, keepdims=True)).astype(np.float32)
    k1 = np.stack([rng.uniform(20, 600, 200), rng.uniform(20, 440, 200)], 1).astype(np.float32)
    k2 = k1[perm] + np.float32(5.0)
    n, m, prev = mo.search_for_initialization(D1, k1, D2, k2, 640, 480, k1.copy())
    inv = np.argsort(perm)
    assert n > 150
    good = m >= 0
    assert np.array_equal(m[good], inv[good])
    assert np.allclose(prev[good], k2[m[good]])
    assert n == int(good.sum())
