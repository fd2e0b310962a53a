#!/usr/bin/env python
"""Probe: time xfb_match_frame_pairs_device for growing pair counts on one extracted VGA batch (debug aid)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from xfeatslam_b200.capi import XFeatB200
from xfeatslam_b200.frames import synthetic_frame
B, H, W, K = 8, 480, 640, 4096
dev = torch.device("cuda", 0)
ctx = XFeatB200(max_h=H, max_w=W, max_batch=B, max_topk=K)
fr = torch.from_numpy(np.stack([synthetic_frame(7000 + i) for i in range(B)])).to(dev)
nv = torch.zeros(B, dtype=torch.int32, device=dev); xy = torch.zeros(B, K, 2, device=dev); sc = torch.zeros(B, K, device=dev); ds = torch.zeros(B, K, 64, device=dev)
ctx.extract_ptrs(fr.data_ptr(), B, H * W, H, W, W, K, 0.05, nv.data_ptr(), xy.data_ptr(), sc.data_ptr(), ds.data_ptr(), device=True)
torch.cuda.synchronize(); print("extract ok, n_valid", nv.tolist(), flush=True)
for npairs in (1, 2, 4, 8):
    pairs = np.array([[i, (i - 1) % B] for i in range(npairs)], np.int32)
    m = [torch.zeros(npairs, K, dtype=torch.int32, device=dev) for _ in range(5)]
    for which in ("fwd", "all"):
        ptrs = [t.data_ptr() for t in m] if which == "all" else [m[0].data_ptr(), m[1].data_ptr(), m[2].data_ptr(), 0, 0]
        t0 = time.time()
        ctx.match_frame_pairs(pairs, 2 ** 31 - 1, ptrs, device=True)
        torch.cuda.synchronize()
        print("pairs %d %s: %.1f ms, matched %d" % (npairs, which, 1e3 * (time.time() - t0), int((m[0] >= 0).sum())), flush=True)
ctx.close()
