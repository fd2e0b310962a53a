"""Parity of the CUDA extract path (through the C-ABI) with the oracle -- runs on the B200 box.

Float stages: tolerance stated per test (target from BASELINE.json: descriptors within 1e-4).
Discrete stages (NMS, score, top-k order) are checked bit-exact on oracle/reference-provided dense
maps, where 1-ulp differences in the CNN cannot move a threshold (SURVEY.md hard part (c))."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import xfeat_oracle as xo
from xfeatslam_b200.frames import synthetic_frame, synthetic_frames

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
BASIC = ["block1.0", "block1.1", "block1.2", "block1.3", "block2.0", "block2.1", "block3.0", "block3.1", "block3.2", "block4.0", "block4.1",
         "block4.2", "block5.0", "block5.1", "block5.2", "block5.3", "block_fusion.0", "block_fusion.1", "heatmap_head.0",
         "heatmap_head.1", "keypoint_head.0", "keypoint_head.1", "keypoint_head.2"]
LAYER_TOL = 6e-4      # abs, raw conv outputs (values up to ~40); the 3xTF32 tensor-core split is good to ~1e-5 relative
DESC_TOL = 1e-4       # BASELINE.json north_star
SCORE_TOL = 5e-5


def nhwc(t):
    return t[0].permute(1, 2, 0).contiguous().numpy()


def gold(name):
    z = np.load(GOLD / (name + ".npz"))
    d = {k.replace("__", "."): z[k] for k in z.files}
    idx, H, W, nfeat, l0, l1 = [int(v) for v in d.pop("meta")]
    return d, synthetic_frame(idx, H, W), nfeat, (l0, l1)


def match_sets(out, kp):
    n = int(out["n_valid"])
    got = {(int(x), int(y)): j for j, (x, y) in enumerate(out["kpts"][:n])}
    common = [(got[(int(x), int(y))], j) for j, (x, y) in enumerate(kp) if (int(x), int(y)) in got]
    gi = np.array([c[0] for c in common], int); oi = np.array([c[1] for c in common], int)
    return n, gi, oi


@pytest.mark.parametrize("shape", [(64, 96), (96, 128)])
def test_every_layer_against_oracle(xfb_small, weights, shape):
    H, W = shape
    frame = synthetic_frame(21, H, W)
    keep = {}
    xo.detect_and_compute(frame, weights, 256, keep)
    xfb_small.extract(frame, 256)
    np.testing.assert_allclose(xfb_small.debug_read("xn")[..., 0], keep["xn"][0, 0].numpy(), atol=2e-5, rtol=0)
    for L in BASIC:
        np.testing.assert_allclose(xfb_small.debug_read(L), nhwc(keep[L + ".conv"]), atol=LAYER_TOL, rtol=0, err_msg=L)
        conv = keep[L + ".conv"][0].double()
        mean, rstd = xfb_small.debug_stats(L)
        np.testing.assert_allclose(mean, conv.mean(dim=(1, 2)).numpy(), atol=2e-4, rtol=0, err_msg=L + " mean")
        want_rstd = 1.0 / np.sqrt(conv.var(dim=(1, 2), unbiased=False).numpy() + 1e-5)
        # 1/sqrt(var + eps) amplifies absolute errors of the raw conv output when a channel's variance is tiny
        # (a handful of pixels at 1/32 resolution): compare the normalised values instead of demanding a tight rstd
        np.testing.assert_allclose(rstd, want_rstd, rtol=5e-3, atol=0, err_msg=L + " rstd")
    np.testing.assert_allclose(xfb_small.debug_read("pyramid_sum"), nhwc(keep["pyramid_sum"]), atol=LAYER_TOL, rtol=0)
    np.testing.assert_allclose(xfb_small.debug_read("feats"), nhwc(keep["feats"]), atol=LAYER_TOL, rtol=0)
    np.testing.assert_allclose(xfb_small.debug_read("H1")[..., 0], keep["H1"][0, 0].numpy(), atol=1e-4, rtol=0)
    np.testing.assert_allclose(xfb_small.debug_read("K1h")[..., 0], keep["K1h"][0, 0].numpy(), atol=1e-4, rtol=0)


@pytest.mark.parametrize("name", ["small_64x96", "resize_100x140", "mono_96x128"])
def test_dense_maps_against_reference_golden(xfb_small, name):
    """Against outputs of the reference binary itself (tests/golden)."""
    g, frame, nfeat, lap = gold(name)
    xfb_small.extract(frame, nfeat)
    np.testing.assert_allclose(xfb_small.debug_read("x_pre")[..., 0], g["x_pre"][0, 0], atol=1e-6, rtol=0)
    np.testing.assert_allclose(xfb_small.debug_read("xn")[..., 0], g["xn"][0, 0], atol=2e-5, rtol=0)
    np.testing.assert_allclose(xfb_small.debug_read("feats"), np.transpose(g["feats"][0], (1, 2, 0)), atol=LAYER_TOL, rtol=0)
    np.testing.assert_allclose(xfb_small.debug_read("H1")[..., 0], g["H1"][0, 0], atol=1e-4, rtol=0)
    np.testing.assert_allclose(xfb_small.debug_read("K1h")[..., 0], g["K1h"][0, 0], atol=1e-4, rtol=0)


@pytest.mark.parametrize("name", ["small_64x96", "resize_100x140", "mono_96x128"])
def test_discrete_stages_bit_exact_on_reference_maps(xfb_small, name):
    """NMS + score + top-k order on the REFERENCE's dense maps: keypoints and scores bit-exact."""
    g, frame, nfeat, lap = gold(name)
    post = xfb_small.debug_post(np.transpose(g["feats"][0], (1, 2, 0)), g["H1"][0, 0], g["K1h"][0, 0], nfeat)
    mk = torch.from_numpy(g["nms_kpts"].astype(np.int64)); sc = torch.from_numpy(g["scores_all"])
    W = g["K1h"].shape[-1]
    order = xo.canonical_order(sc, mk, W)[:nfeat]
    kp_ref = g["nms_kpts"][0][order].astype(np.int64); sc_ref = g["scores_all"][0][order]
    valid = sc_ref > 0
    kp_ref, sc_ref = kp_ref[valid], sc_ref[valid]
    n = post["n_valid"]
    assert n == len(kp_ref)
    assert np.array_equal(post["kpts"][:n].astype(np.int64), kp_ref)
    assert np.array_equal(post["scores"][:n], sc_ref)                      # bit-exact
    lin = {(int(x), int(y)): i for i, (x, y) in enumerate(g["nms_kpts"][0])}
    rows = [lin[(int(x), int(y))] for x, y in kp_ref]
    np.testing.assert_allclose(post["desc"][:n], g["desc_all"][0][rows], atol=1e-6, rtol=0)
    assert np.all(post["kpts"][n:] == 0) and np.all(post["desc"][n:] == 0) and np.all(post["scores"][n:] == 0)


def test_nms_plateau_lists_every_pixel():
    """A plateau above the threshold: every pixel of it equals its 5x5 maximum, so a whole 64 x 16 NMS tile (1024 pixels) lands on the
    kernel's candidate list (its scoring loop then runs more than one round).  Synthetic dense maps through the same debug entry as
    above, expected values from the oracle's restatement of XFextractor::NMS / the score (src/XFextractor.cc:219-248, :277-282)."""
    rng = np.random.RandomState(5)
    H, W, nfeat = 64, 96, 4096
    K1h = (rng.rand(H, W).astype(np.float32) * 0.04).astype(np.float32)          # background below the 0.05 threshold
    K1h[8:56, 16:88] = np.float32(0.3)                                           # the plateau: 48 x 72 pixels
    K1h[3, 5] = np.float32(0.9)                                                  # and an ordinary isolated peak
    H1 = rng.rand(H // 8, W // 8).astype(np.float32)
    feats = rng.randn(H // 8, W // 8, 64).astype(np.float32)
    from xfeatslam_b200.capi import XFeatB200
    ctx = XFeatB200(max_h=H, max_w=W, max_batch=1, max_topk=nfeat)
    post = ctx.debug_post(feats, H1, K1h, nfeat)
    ctx.close()
    K1h_t = torch.from_numpy(K1h)[None, None]; H1_t = torch.from_numpy(H1)[None, None]
    mk = xo.nms(K1h_t, 0.05, 5)
    sn = xo.interpolate_sparse2d(K1h_t, mk, H, W, "nearest"); sb = xo.interpolate_sparse2d(H1_t, mk, H, W, "bilinear")
    sc = (sn * sb).squeeze(-1).masked_fill(torch.all(mk == 0, -1), -1)
    assert mk.shape[1] >= 48 * 72
    order = xo.canonical_order(sc, mk, W)[:nfeat]
    kp_ref = mk[0].numpy()[order]; sc_ref = sc[0].numpy()[order]
    valid = sc_ref > 0
    kp_ref, sc_ref = kp_ref[valid], sc_ref[valid]
    n = post["n_valid"]
    assert n == len(kp_ref) and n > 3000
    assert np.array_equal(post["kpts"][:n].astype(np.int64), kp_ref)
    assert np.array_equal(post["scores"][:n], sc_ref)


@pytest.mark.parametrize("name,tol_common", [("vga_top4096", 0.99), ("vga_top1000", 0.98), ("hd720_top1000", 0.98)])
def test_end_to_end_against_reference_golden(name, tol_common):
    from xfeatslam_b200.capi import XFeatB200
    g, frame, nfeat, lap = gold(name)
    ctx = XFeatB200(max_h=frame.shape[0], max_w=frame.shape[1], max_batch=1, max_topk=nfeat)
    out = ctx.extract(frame, nfeat)
    gk, gd = g["out_keypoints"], g["out_descriptors"]
    valid = gk[:, 2] > 0
    n, gi, oi = match_sets(out, gk[valid][:, :2])
    assert n == int(valid.sum())
    assert len(gi) >= tol_common * n                                         # set differs only at the k-th score boundary
    np.testing.assert_allclose(out["scores"][gi], gk[valid][oi, 2], atol=SCORE_TOL, rtol=0)
    np.testing.assert_allclose(out["desc"][gi], gd[valid][oi], atol=DESC_TOL, rtol=0)
    # sorted by score, unit norm
    s = out["scores"][:n]
    assert np.all(s[:-1] >= s[1:])
    np.testing.assert_allclose(np.linalg.norm(out["desc"][:n], axis=1), 1.0, atol=1e-5)
    ctx.close()


@pytest.mark.parametrize("name", ["vga_top4096", "vga_top1000", "hd720_top1000"])
def test_keypoint_set_differs_only_at_the_kth_score_boundary(name):
    """The claim behind the '>= 98 % common keypoints' tolerances, asserted: a keypoint of the reference's top-k that is
    missing from ours (or vice versa) sits AT the k-th score boundary -- its reference score is within 2 * SCORE_TOL of the
    reference's k-th score -- or is one of a handful of NMS flips (a 5x5 maximum / the 0.05 threshold decided by < 1e-5)."""
    from xfeatslam_b200.capi import XFeatB200
    g, frame, nfeat, lap = gold(name)
    ctx = XFeatB200(max_h=frame.shape[0], max_w=frame.shape[1], max_batch=1, max_topk=8192)
    full = ctx.extract(frame, 8192)                                          # our candidates in score order (all of them at VGA)
    n_full = int(full["n_valid"])
    ours_all = {(int(x), int(y)): float(s) for (x, y), s in zip(full["kpts"][:n_full], full["scores"][:n_full])}
    ours_top = {(int(x), int(y)) for x, y in full["kpts"][:nfeat]}           # top-k is a prefix (test_topk_8192_returns_every_candidate)
    ref_all = {(int(x), int(y)): float(s) for (x, y), s in zip(g["nms_kpts"][0], g["scores_all"][0]) if s > 0}
    gk = g["out_keypoints"]
    gk = gk[gk[:, 2] > 0]
    ref_top = {(int(x), int(y)) for x, y in gk[:, :2]}
    assert len(ref_top) == nfeat and len(ours_top) == nfeat
    kth_ref = float(gk[:, 2].min())
    flips = 0
    for p in ref_top - ours_top:
        if p not in ours_all and n_full < 8192:
            flips += 1
            continue
        assert ref_all[p] - kth_ref <= 2 * SCORE_TOL, ("reference keypoint missing although well above the cut-off", p, ref_all[p], kth_ref)
    for p in ours_top - ref_top:
        if p not in ref_all:
            flips += 1
            continue
        assert kth_ref - ref_all[p] <= 2 * SCORE_TOL, ("extra keypoint although well below the reference's cut-off", p, ref_all[p], kth_ref)
    print("%s: |ref - ours| = %d, |ours - ref| = %d, NMS flips = %d" % (name, len(ref_top - ours_top), len(ours_top - ref_top), flips))
    assert flips <= 3
    common = sorted(ours_top & ref_top)
    np.testing.assert_allclose([ours_all[p] for p in common], [ref_all[p] for p in common], atol=SCORE_TOL, rtol=0)
    ctx.close()


def test_block1_fusion_equals_two_kernels(weights, monkeypatch):
    """block1.0 recomputed inside block1.1 (XFB_B1_FUSE=1, an A/B option: measured slower than the two kernels) against the
    default two-kernel form: the recomputation uses the same accumulation order, so every output is bit-identical."""
    from xfeatslam_b200.capi import XFeatB200
    frames = synthetic_frames(33, 2, 96, 128)
    plain = XFeatB200(max_h=96, max_w=128, max_batch=2, max_topk=256)
    op = plain.extract(frames, 256)
    monkeypatch.setenv("XFB_B1_FUSE", "1")
    fused = XFeatB200(max_h=96, max_w=128, max_batch=2, max_topk=256)
    of = fused.extract(frames, 256)
    for k in ("n_valid", "kpts", "scores", "desc"):
        assert np.array_equal(of[k], op[k]), k
    for L in ("block1.1", "block1.3", "block_fusion.1"):
        assert np.array_equal(fused.debug_read(L, frame=1), plain.debug_read(L, frame=1)), L
    keep = {}
    xo.detect_and_compute(frames[1], weights, 256, keep)
    np.testing.assert_allclose(plain.debug_read("block1.0", frame=1), nhwc(keep["block1.0.conv"]), atol=LAYER_TOL, rtol=0)
    m_f, r_f = fused.debug_stats("block1.0", frame=1)
    m_p, r_p = plain.debug_stats("block1.0", frame=1)
    assert np.array_equal(m_f, m_p) and np.array_equal(r_f, r_p)
    fused.close(); plain.close()


def test_batch_equals_single_frames(xfb_vga):
    """Per-frame BatchNorm statistics: a batch launch equals B batch-1 runs, bit for bit."""
    frames = synthetic_frames(40, 3, 480, 640)
    ob = xfb_vga.extract(frames, 2048)
    for i in range(3):
        o1 = xfb_vga.extract(frames[i], 2048)
        for k in ("kpts", "scores", "desc"):
            assert np.array_equal(ob[k][i], o1[k]), (i, k)
        assert int(ob["n_valid"][i]) == int(o1["n_valid"])


def test_deterministic_run_to_run(xfb_vga):
    f = synthetic_frame(50)
    a = xfb_vga.extract(f, 4096)
    b = xfb_vga.extract(f, 4096)
    for k in ("kpts", "scores", "desc"):
        assert np.array_equal(a[k], b[k])


def test_full_size_properties(xfb_vga, weights):
    f = synthetic_frame(60)
    out = xfb_vga.extract(f, 4096)
    n = int(out["n_valid"])
    assert n == 4096
    xy = out["kpts"][:n]
    assert np.all(xy == np.rint(xy)) and xy[:, 0].max() < 639 and xy[:, 1].max() < 479 and xy.min() >= 0
    assert len({(int(x), int(y)) for x, y in xy}) == n                      # unique pixels
    # 5x5 NMS: no two keypoints closer than 3 px in Chebyshev distance unless the heat map ties
    kp, sc, ds = xo.detect_and_compute(f, weights, 4096)
    n2, gi, oi = match_sets(out, kp)
    assert len(gi) >= 0.99 * n
    np.testing.assert_allclose(out["desc"][gi], ds[oi], atol=DESC_TOL, rtol=0)


def test_edge_cases(xfb_small):
    from xfeatslam_b200.capi import XFBError
    # constant image: InstanceNorm of a constant is 0 everywhere -> no crash, deterministic output
    flat = np.full((64, 96), 77, np.uint8)
    o = xfb_small.extract(flat, 64)
    assert 0 <= int(o["n_valid"]) <= 64
    # smallest legal frame: two pixels at 1/32 resolution (with one, the reference's train-mode BatchNorm throws
    # "Expected more than 1 value per channel when training" -- mirrored as an error code)
    o = xfb_small.extract(synthetic_frame(5, 32, 64), 16)
    assert 0 <= int(o["n_valid"]) <= 16
    with pytest.raises(XFBError):
        xfb_small.extract(synthetic_frame(5, 32, 32), 16)
    with pytest.raises(XFBError):
        xfb_small.extract(synthetic_frame(5, 40, 50), 16)
    # topk larger than the number of candidates -> padded with zeros
    o = xfb_small.extract(synthetic_frame(6, 64, 64), 1024)
    n = int(o["n_valid"])
    assert n < 1024 and np.all(o["scores"][n:] == 0) and np.all(o["desc"][n:] == 0)
    assert n == xfb_small.candidates(0)
    # too large / too small / bad topk -> error code, not a crash
    with pytest.raises(XFBError):
        xfb_small.extract(np.zeros((256, 256), np.uint8), 16)
    with pytest.raises(XFBError):
        xfb_small.extract(np.zeros((16, 64), np.uint8), 16)
    with pytest.raises(XFBError):
        xfb_small.extract(synthetic_frame(1, 64, 64), 5000)


def test_row_stride_is_honoured(xfb_small):
    import ctypes
    f = synthetic_frame(8, 64, 96)
    padded = np.zeros((64, 128), np.uint8)
    padded[:, :96] = f
    topk = 128
    nv = np.zeros(1, np.int32); xy = np.zeros((topk, 2), np.float32); sc = np.zeros(topk, np.float32); ds = np.zeros((topk, 64), np.float32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = xfb_small.lib.xfb_extract(xfb_small.h, p(padded), 64, 96, 128, topk, 0.05, p(nv), p(xy), p(sc), p(ds))
    assert rc == 0
    ref = xfb_small.extract(f, topk)
    assert np.array_equal(xy, ref["kpts"]) and np.array_equal(ds, ref["desc"])


def test_pipelined_submit_equals_synchronous_calls(xfb_vga):
    """xfb_submit / xfb_wait (three streams, two slots in flight) returns exactly what the synchronous
    entry points return."""
    import torch
    frames = [synthetic_frames(70 + 3 * i, 3, 480, 640) for i in range(4)]
    topk = 1024
    pairs = np.array([[1, 0], [2, 1], [0, 2]], np.int32)
    want = []
    for f in frames:
        o = xfb_vga.extract(f, topk)
        m = [np.zeros((3, topk), np.int32) for _ in range(5)]
        xfb_vga.match_frame_pairs(pairs, 2 ** 31 - 1, [x.ctypes.data for x in m])
        want.append((o, m))
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    outs = [{"nv": torch.zeros(3, dtype=torch.int32).pin_memory(), "xy": torch.zeros(3, topk, 2).pin_memory(), "sc": torch.zeros(3, topk).pin_memory(),
             "ds": torch.zeros(3, topk, 64).pin_memory(), "m": [torch.zeros(3, topk, dtype=torch.int32).pin_memory() for _ in range(5)]} for _ in range(2)]
    got = []
    for i in range(4):
        slot = i % 2
        if i >= 2:
            xfb_vga.wait(slot)
            o = outs[slot]
            got.append(({"n_valid": o["nv"].numpy().copy(), "kpts": o["xy"].numpy().copy(), "scores": o["sc"].numpy().copy(), "desc": o["ds"].numpy().copy()},
                        [t.numpy().copy() for t in o["m"]]))
        o = outs[slot]
        xfb_vga.submit(slot, pinned[i].data_ptr(), 3, 480 * 640, 480, 640, 640, topk, 0.05, o["nv"].data_ptr(), o["xy"].data_ptr(), o["sc"].data_ptr(),
                       o["ds"].data_ptr(), pairs=pairs, match_ptrs=[t.data_ptr() for t in o["m"]])
    for slot in (0, 1):
        xfb_vga.wait(slot)
        o = outs[slot]
        got.append(({"n_valid": o["nv"].numpy().copy(), "kpts": o["xy"].numpy().copy(), "scores": o["sc"].numpy().copy(), "desc": o["ds"].numpy().copy()},
                    [t.numpy().copy() for t in o["m"]]))
    assert len(got) == 4
    for (wo, wm), (go, gm) in zip(want, got):
        for k in ("n_valid", "kpts", "scores", "desc"):
            assert np.array_equal(wo[k], go[k]), k
        for a, b in zip(wm, gm):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", [(40, 70), (77, 131), (128, 160), (96, 160)])
def test_odd_sizes_against_oracle(xfb_small, weights, shape):
    """Sizes that are not multiples of 32 go through the bilinear pre-resize (src/XFextractor.cc:182-202);
    keypoints stay in the resized frame (SURVEY finding 2)."""
    H, W = shape
    frame = synthetic_frame(31 + H, H, W)
    kp, sc, ds = xo.detect_and_compute(frame, weights, 512)
    out = xfb_small.extract(frame, 512)
    n, gi, oi = match_sets(out, kp)
    assert abs(n - len(kp)) <= max(2, len(kp) // 50)
    assert len(gi) >= 0.95 * len(kp)
    if len(gi):
        # a 1 x 2-pixel map at 1/32 resolution (40 x 70 -> 32 x 64) puts TWO samples into block5's train-mode BatchNorm: the
        # normalised values are +-d / sqrt(d^2 + 4e-5), which amplifies fp32-level differences of the raw conv output whenever a
        # channel's two samples nearly coincide -- the reference itself is ill-conditioned there
        tiny = (H // 32) * (W // 32) <= 2
        np.testing.assert_allclose(out["scores"][gi], sc[oi], atol=3 * SCORE_TOL if tiny else SCORE_TOL, rtol=0)
        np.testing.assert_allclose(out["desc"][gi], ds[oi], atol=3 * DESC_TOL if tiny else DESC_TOL, rtol=0)
        assert out["kpts"][:n, 0].max() < (W // 32) * 32 and out["kpts"][:n, 1].max() < (H // 32) * 32


def test_topk_8192_returns_every_candidate():
    from xfeatslam_b200.capi import XFeatB200
    ctx = XFeatB200(max_h=480, max_w=640, max_batch=1, max_topk=8192)
    f = synthetic_frame(91)
    out = ctx.extract(f, 8192)
    n = int(out["n_valid"])
    assert n == ctx.candidates(0) and 4096 < n < 8192          # ~7000 NMS survivors with score > 0 at VGA
    s = out["scores"][:n]
    assert np.all(s[:-1] >= s[1:]) and s[-1] > 0
    assert len({(int(x), int(y)) for x, y in out["kpts"][:n]}) == n
    top = ctx.extract(f, 1000)
    assert np.array_equal(top["kpts"], out["kpts"][:1000]) and np.array_equal(top["desc"], out["desc"][:1000])   # top-k is a prefix
    ctx.close()


def test_hd720_batch_equals_singles():
    from xfeatslam_b200.capi import XFeatB200
    ctx = XFeatB200(max_h=720, max_w=1280, max_batch=2, max_topk=2000)
    frames = synthetic_frames(120, 2, 720, 1280)
    ob = ctx.extract(frames, 2000)
    for i in range(2):
        o1 = ctx.extract(frames[i], 2000)
        for k in ("kpts", "scores", "desc"):
            assert np.array_equal(ob[k][i], o1[k])
        assert o1["kpts"][:, 1].max() < 704                     # 720 -> 704 rows internally, never rescaled
    ctx.close()
