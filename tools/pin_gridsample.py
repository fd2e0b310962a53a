#!/usr/bin/env python
"""Finds the exact rounding sequence of ATen's fp32 CPU grid_sample (bilinear, zeros, align_corners=False)
as called by the reference's InterpolateSparse2d (src/XFeat.cc:181-210), by emulating candidate op orders in
numpy (fma emulated in float64) and comparing bit-for-bit with torch.  Result (torch 2.11, AVX2/AVX512 build):

    src = fma(g + 1, S_map / 2, -0.5)
    out = fma(se_v, se, fma(sw_v, sw, fma(ne_v, ne, nw_v * nw)))

csrc/post.cu implements exactly this with __fmaf_rn / __fmul_rn, which is why the GPU scores are
bit-identical to the reference on identical dense maps.  TEST TOOLING ONLY."""
import itertools
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

f32, f64 = np.float32, np.float64


def fma(a, b, c):
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)


def main():
    rng = np.random.RandomState(0)
    H, W, mh, mw = 480, 640, 60, 80
    hm = rng.rand(mh, mw).astype(f32)
    xs = rng.randint(0, W, 20000).astype(f32); ys = rng.randint(0, H, 20000).astype(f32)
    pos = torch.from_numpy(np.stack([xs, ys], 1).astype(np.int64))[None]
    size = torch.tensor([W - 1, H - 1], dtype=pos.dtype)
    grid = (2.0 * (pos / size) - 1.0).unsqueeze(-2).to(torch.float32)
    ref = F.grid_sample(torch.from_numpy(hm)[None, None], grid, mode="bilinear", padding_mode="zeros", align_corners=False)[0, 0, :, 0].numpy()

    def src(p, full, m, fused):
        g = f32(2.0) * (p / f32(full - 1)) - f32(1.0)
        a = g + f32(1.0)
        return fma(a, np.full_like(p, m / 2), np.full_like(p, -0.5)) if fused else a * f32(m / 2) - f32(0.5)

    def gather(yy, xx):
        yy = yy.astype(int); xx = xx.astype(int)
        ok = (yy >= 0) & (yy < mh) & (xx >= 0) & (xx < mw)
        return np.where(ok, hm[np.clip(yy, 0, mh - 1), np.clip(xx, 0, mw - 1)], f32(0)).astype(f32)

    for fused in (False, True):
        sx, sy = src(xs, W, mw, fused), src(ys, H, mh, fused)
        x0, y0 = np.floor(sx), np.floor(sy)
        w = sx - x0; e = f32(1) - w; n = sy - y0; s = f32(1) - n
        v = [gather(y0, x0), gather(y0, x0 + 1), gather(y0 + 1, x0), gather(y0 + 1, x0 + 1)]
        wt = [s * e, s * w, n * e, n * w]
        cands = {
            "unfused": ((v[0] * wt[0] + v[1] * wt[1]) + v[2] * wt[2]) + v[3] * wt[3],
            "fma_chain": fma(v[3], wt[3], fma(v[2], wt[2], fma(v[1], wt[1], v[0] * wt[0]))),
            "fma_chain_rev": fma(v[0], wt[0], fma(v[1], wt[1], fma(v[2], wt[2], v[3] * wt[3]))),
        }
        for name, c in cands.items():
            print("src %-8s blend %-14s mismatches %6d / %d" % ("fma" if fused else "unfused", name, int((c != ref).sum()), len(ref)))


if __name__ == "__main__":
    sys.exit(main())
