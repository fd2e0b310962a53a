"""Bit-exact parity of the CUDA matcher (through the C-ABI) with oracle/matcher_oracle.c."""
import numpy as np
import pytest

from oracle import matcher_oracle as mo

pytestmark = pytest.mark.gpu
INT_MAX = 2 ** 31 - 1


def unit(rng, n):
    a = rng.randn(n, 64).astype(np.float32)
    return a / np.linalg.norm(a, axis=1, keepdims=True)


def related(rng, A, n2, noise):
    idx = rng.randint(0, A.shape[0], n2)
    B = A[idx] + noise * rng.randn(n2, 64).astype(np.float32)
    return (B / np.linalg.norm(B, axis=1, keepdims=True)).astype(np.float32)


@pytest.fixture(scope="module")
def ctx():
    from xfeatslam_b200.capi import XFeatB200
    c = XFeatB200(max_h=64, max_w=64, max_batch=1, max_topk=16)
    yield c
    c.close()


@pytest.mark.parametrize("n1,n2", [(1, 1), (63, 65), (64, 64), (257, 130), (1000, 777)])
def test_distance_matrix_bit_exact(ctx, n1, n2):
    rng = np.random.RandomState(n1 * 1000 + n2)
    A = unit(rng, n1); B = related(rng, A, n2, 0.08)
    assert np.array_equal(ctx.distance_matrix(A, B), mo.distance_matrix(A, B))


@pytest.mark.parametrize("init", [INT_MAX, 256])
@pytest.mark.parametrize("n1,n2", [(1, 5), (100, 64), (513, 1000), (1500, 1400)])
def test_match_bit_exact(ctx, n1, n2, init):
    rng = np.random.RandomState(n1 + 7 * n2)
    A = unit(rng, n1); B = related(rng, A, n2, 0.05)
    if n2 > 20:
        B[11] = B[3]                      # exact duplicate column: lowest index must win
        A[0] = 0                          # phantom (all-zero) descriptor rows, SURVEY 8a M1
        B[7] = 0; B[8] = 0
    got = ctx.match(A, B, init=init)
    want = mo.bruteforce(A, B, init=init)
    for name, g, w in zip(("best_idx", "best_dist", "second_dist", "rev_idx", "rev_dist"), got, want):
        assert np.array_equal(g, w), name


def test_match_with_group_gating(ctx):
    rng = np.random.RandomState(5)
    A = unit(rng, 700); B = related(rng, A, 650, 0.05)
    ga = rng.randint(0, 30, 700); gb = rng.randint(0, 30, 650)
    ga[5] = 99                            # group with no partner -> idx -1, dist = init
    got = ctx.match(A, B, ga, gb, init=256)
    want = mo.bruteforce(A, B, ga, gb, init=256)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert got[0][5] == -1 and got[1][5] == 256


def test_match_empty_sides(ctx):
    A = unit(np.random.RandomState(1), 10)
    bi, bd, sd, ri, rd = ctx.match(A, np.zeros((0, 64), np.float32))
    assert np.all(bi == -1) and np.all(bd == INT_MAX) and len(ri) == 0
    bi, bd, sd, ri, rd = ctx.match(np.zeros((0, 64), np.float32), A)
    assert len(bi) == 0 and np.all(ri == -1)


def test_full_size_4096_properties(ctx):
    """BASELINE config 3 size: 4096 x 4096.  The oracle would need ~1e9 flops in C (a few s): run it on
    a row sample, and check size-independent properties on everything."""
    rng = np.random.RandomState(9)
    A = unit(rng, 4096)
    perm = rng.permutation(4096)
    B = A[perm] + 0.03 * rng.randn(4096, 64).astype(np.float32)
    B = (B / np.linalg.norm(B, axis=1, keepdims=True)).astype(np.float32)
    bi, bd, sd, ri, rd = ctx.match(A, B)
    inv = np.argsort(perm)
    assert (bi == inv).mean() > 0.999                      # recovers the permutation
    assert np.all(bd <= sd)
    mutual = ri[bi] == np.arange(4096)
    assert mutual.mean() > 0.999
    # symmetry: swapping the operands swaps the forward / reverse outputs
    bi2, bd2, sd2, ri2, rd2 = ctx.match(B, A)
    assert np.array_equal(ri2, bi) and np.array_equal(rd2, bd) and np.array_equal(bi2, ri) and np.array_equal(bd2, rd)
    rows = rng.choice(4096, 64, replace=False)
    w = mo.bruteforce(A[rows], B)
    assert np.array_equal(bi[rows], w[0]) and np.array_equal(bd[rows], w[1]) and np.array_equal(sd[rows], w[2])


def test_tensor_core_estimate_error_bound(ctx):
    """The matcher filters on a tensor-core estimate t of 512*d (bf16 split a1.b1 + a1.b2 + a2.b1); the filter is
    sound while |t - 512*float(S)| < match_eps(|a|^2, |b|^2) = 0.055 for unit rows (csrc/match_tc.cu).  Measure the actual maximum."""
    rng = np.random.RandomState(13)
    A = unit(rng, 1024)
    B = np.concatenate([related(rng, A, 512, 0.02), related(rng, A, 256, 0.3), unit(rng, 256)]).astype(np.float32)
    B[3] = 0                                               # phantom row
    B[4] = A[0]                                            # exact duplicate -> distance 0
    err = ctx.match_error(A, B)
    assert err < 0.02, err


@pytest.mark.parametrize("scale_a,scale_b", [(3.0, 0.25), (40.0, 40.0), (1e-3, 1.0), (400.0, 1.0)])
def test_non_unit_norm_rows_bit_exact(ctx, scale_a, scale_b):
    """The C-ABI takes arbitrary finite fp32 rows: the tensor-core shortcuts carry PER-PAIR error bounds that scale with
    |a||b| (csrc/match_tc.cu match_eps, csrc/match_stream.cu e / `wild`), so results stay bit-exact off the unit sphere
    (rows with |x|^2 >= 1e5 leave the fp16 filter's range and are verified exhaustively)."""
    rng = np.random.RandomState(int(scale_a * 1000 + scale_b * 10) % 2 ** 31)
    A = (unit(rng, 300) * scale_a * rng.uniform(0.5, 1.5, (300, 1))).astype(np.float32)
    B = (related(rng, unit(rng, 300), 260, 0.05) * scale_b * rng.uniform(0.5, 1.5, (260, 1))).astype(np.float32)
    B[5] = A[7]                                            # an exact duplicate across the sets -> distance 0
    # distances stay inside the int range: ||a-b||^2 * 512 < 2^31
    assert (np.linalg.norm(A, axis=1).max() + np.linalg.norm(B, axis=1).max()) ** 2 * 512 < 2 ** 31
    assert np.array_equal(ctx.distance_matrix(A, B), mo.distance_matrix(A, B))
    for init in (INT_MAX, 256):
        got = ctx.match(A, B, init=init)
        want = mo.bruteforce(A, B, init=init)
        for name, g, w in zip(("best_idx", "best_dist", "second_dist", "rev_idx", "rev_dist"), got, want):
            assert np.array_equal(g, w), (name, init)
