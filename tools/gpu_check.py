#!/usr/bin/env python
"""Layer-by-layer parity report of the CUDA path against the torch-CPU oracle (runs on the GPU box).
Prints max-abs errors for every intermediate; used for bring-up, the asserting versions live in tests/."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle import xfeat_oracle as xo  # noqa: E402
from xfeatslam_b200.capi import XFeatB200  # noqa: E402
from xfeatslam_b200.frames import synthetic_frame  # noqa: E402

LAYERS = ["block1.0", "block1.1", "block1.2", "block1.3", "block2.0", "block2.1", "block3.0", "block3.1", "block3.2", "block4.0",
          "block4.1", "block4.2", "block5.0", "block5.1", "block5.2", "block5.3", "block_fusion.0", "block_fusion.1",
          "heatmap_head.0", "heatmap_head.1", "keypoint_head.0", "keypoint_head.1", "keypoint_head.2"]


def nhwc(t):
    return t[0].permute(1, 2, 0).contiguous().numpy()


def report(name, got, want):
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    if got.shape != want.shape:
        print("%-22s SHAPE MISMATCH %s vs %s" % (name, got.shape, want.shape)); return
    err = np.abs(got - want)
    print("%-22s shape %-16s maxabs %.3e  mean %.3e  ref-absmax %.3e" % (name, got.shape, err.max(), err.mean(), np.abs(want).max()))


def main():
    H, W, topk = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (480, 640, 4096)
    idx = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    wts = xo.load_weights()
    frame = synthetic_frame(idx, H, W)
    keep = {}
    t0 = time.time()
    kp, sc, ds = xo.detect_and_compute(frame, wts, topk, keep)
    print("oracle: %.2fs, N_nms=%d valid=%d" % (time.time() - t0, keep["nms_kpts"].shape[1], len(kp)))
    ctx = XFeatB200(max_h=H, max_w=W, max_batch=2, max_topk=topk)
    out = ctx.extract(frame, topk)
    print("gpu: n_valid=%d candidates=%d launches=%d" % (out["n_valid"], ctx.candidates(0), ctx.launch_count()))
    report("x_pre", ctx.debug_read("x_pre")[..., 0], keep["x_pre"][0, 0].numpy())
    report("xn", ctx.debug_read("xn")[..., 0], keep["xn"][0, 0].numpy())
    report("avg4", ctx.debug_read("avg4")[..., 0], keep["skip1.0"][0, 0].numpy())
    for L in LAYERS:
        report(L + ".conv", ctx.debug_read(L), nhwc(keep[L + ".conv"]))
        m, r = ctx.debug_stats(L)
        conv = keep[L + ".conv"][0].double()
        mu = conv.mean(dim=(1, 2)).numpy(); var = conv.var(dim=(1, 2), unbiased=False).numpy()
        report(L + ".mean", m, mu)
        report(L + ".rstd", r, 1.0 / np.sqrt(var + 1e-5))
        if L == "block5.3":
            report("pyramid_sum", ctx.debug_read("pyramid_sum"), nhwc(keep["pyramid_sum"]))
    report("feats", ctx.debug_read("feats"), nhwc(keep["feats"]))
    report("H1", ctx.debug_read("H1")[..., 0], keep["H1"][0, 0].numpy())
    report("K1h", ctx.debug_read("K1h")[..., 0], keep["K1h"][0, 0].numpy())
    # end-to-end keypoints
    n = int(out["n_valid"])
    gk = out["kpts"][:n].astype(np.int64)
    oset = {(int(a), int(b)): i for i, (a, b) in enumerate(kp)}
    common = [(i, oset[(int(a), int(b))]) for i, (a, b) in enumerate(gk) if (int(a), int(b)) in oset]
    print("e2e: gpu %d kpts, oracle %d, common %d" % (n, len(kp), len(common)))
    if common:
        gi = np.array([c[0] for c in common]); oi = np.array([c[1] for c in common])
        print("   score maxabs %.3e   desc maxabs %.3e   same-rank %d" % (np.abs(out["scores"][gi] - sc[oi]).max(),
                                                                          np.abs(out["desc"][gi] - ds[oi]).max(), int((gi == oi).sum())))
    # discrete stages on oracle-provided dense maps
    post = ctx.debug_post(nhwc(keep["feats"]), keep["H1"][0, 0].numpy(), keep["K1h"][0, 0].numpy(), topk)
    n2 = post["n_valid"]
    print("post(oracle maps): n_valid %d vs %d; kpts equal %s; scores equal %s; desc maxabs %.3e" % (
        n2, len(kp), np.array_equal(post["kpts"][:n2].astype(np.int64), kp) if n2 == len(kp) else False,
        np.array_equal(post["scores"][:n2], sc) if n2 == len(kp) else False,
        np.abs(post["desc"][:n2] - ds).max() if n2 == len(kp) else -1))
    # batch == singles
    f2 = np.stack([frame, synthetic_frame(idx + 1, H, W)])
    ob = ctx.extract(f2, topk)
    print("batch[0]==single: kpts %s scores %s desc %s" % (np.array_equal(ob["kpts"][0], out["kpts"]), np.array_equal(ob["scores"][0], out["scores"]),
                                                          np.array_equal(ob["desc"][0], out["desc"])))
    # timing
    import ctypes
    t0 = time.time()
    for _ in range(20):
        ctx.extract(f2, topk)
    print("host-API extract: %.3f ms / frame (batch 2, sync, incl. copies)" % ((time.time() - t0) / 40 * 1e3))




def match_check():
    from oracle import matcher_oracle as mo
    rng = np.random.RandomState(7)
    A = rng.randn(1000, 64).astype(np.float32); A /= np.linalg.norm(A, axis=1, keepdims=True)
    B = (A[rng.permutation(1000)[:900]] + 0.05 * rng.randn(900, 64)).astype(np.float32); B /= np.linalg.norm(B, axis=1, keepdims=True)
    A[17] = 0; B[5] = 0; B[6] = 0
    ctx = XFeatB200(max_h=64, max_w=64, max_batch=1, max_topk=16)
    M = ctx.distance_matrix(A, B)
    Mo = mo.distance_matrix(A, B)
    print("distance matrix equal:", np.array_equal(M, Mo), "mismatches", int((M != Mo).sum()))
    for init in (2 ** 31 - 1, 256):
        got = ctx.match(A, B, init=init)
        want = mo.bruteforce(A, B, init=init)
        print("match init=%d equal:" % init, [bool(np.array_equal(g, w)) for g, w in zip(got, want)])
    ga = rng.randint(0, 20, 1000).astype(np.int32); gb = rng.randint(0, 20, 900).astype(np.int32)
    got = ctx.match(A, B, ga, gb, init=256); want = mo.bruteforce(A, B, ga, gb, init=256)
    print("match grouped equal:", [bool(np.array_equal(g, w)) for g, w in zip(got, want)])


def ab_check(H=480, W=640):
    """tcgen05 conv path vs the FP32 SIMT conv path on the same frame, layer by layer."""
    frame = synthetic_frame(2, H, W)
    ctx = XFeatB200(max_h=H, max_w=W, max_batch=1, max_topk=1024)
    ctx.force_simt(True)
    o_s = ctx.extract(frame, 1024)
    simt = {L: ctx.debug_read(L) for L in LAYERS}
    simt["feats"] = ctx.debug_read("feats")
    ctx.force_simt(False)
    o_t = ctx.extract(frame, 1024)
    for L in LAYERS + ["feats"]:
        report("tc-vs-simt " + L, ctx.debug_read(L), simt[L])
    print("keypoints equal:", np.array_equal(o_s["kpts"], o_t["kpts"]), " desc maxabs %.3e" % np.abs(o_s["desc"] - o_t["desc"]).max())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "match":
        match_check()
    elif len(sys.argv) > 1 and sys.argv[1] == "ab":
        ab_check(*[int(x) for x in sys.argv[2:4]])
    else:
        main()
