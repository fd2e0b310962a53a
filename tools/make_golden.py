#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference itself -- TEST INFRASTRUCTURE ONLY.

Runs oracle/_ref/ref_xfeat (the reference's XFeat.cc + XFextractor.cc compiled unchanged, see
oracle/refbuild/Makefile) on seeded synthetic frames (xfeatslam_b200/frames.py) and stores its
outputs as compact fixtures.  Needs /root/reference (to build oracle/_ref); the fixtures it writes
are what travels to the GPU box.  The reference has no golden vectors of its own.

  small cases : every intermediate tensor of XFextractor::operator() (src/XFextractor.cc:250-357)
  VGA / 720p  : final outputs + the NMS candidate list and scores
"""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle.refdump import read_dump  # noqa: E402
from xfeatslam_b200.frames import synthetic_frame  # noqa: E402

REF = REPO / "oracle" / "_ref" / "ref_xfeat"
OUT = REPO / "tests" / "golden"

# name, frame idx, H, W, nfeatures, lapping, keep-all-intermediates
CASES = [
    ("small_64x96", 3, 64, 96, 256, (0, 0), True),
    ("resize_100x140", 4, 100, 140, 256, (0, 0), True),
    ("mono_96x128", 5, 96, 128, 128, (0, 1000), True),
    ("vga_top4096", 0, 480, 640, 4096, (0, 0), False),
    ("vga_top1000", 1, 480, 640, 1000, (0, 0), False),
    ("hd720_top1000", 2, 720, 1280, 1000, (0, 0), False),
]
FINAL_KEYS = ["nms_kpts", "scores_all", "out_keypoints", "out_descriptors", "out_ret"]


def run_case(name, idx, H, W, nfeat, lap, full):
    frame = synthetic_frame(idx, H, W)
    with tempfile.TemporaryDirectory() as td:
        fp, op = Path(td) / "f.u8", Path(td) / "o.bin"
        frame.tofile(fp)
        subprocess.run([str(REF), "dump", str(fp), str(H), str(W), str(nfeat), str(lap[0]), str(lap[1]), str(op), "4"],
                       check=True, capture_output=True)
        d = read_dump(op)
    keep = {}
    for k, v in d.items():
        if not full and k not in FINAL_KEYS:
            continue
        if k == "nms_kpts":
            v = v.astype(np.int16)
        if k == "out_keypoints":
            v = v[:, :3]
        keep[k.replace(".", "__")] = v
    keep["meta"] = np.array([idx, H, W, nfeat, lap[0], lap[1]], np.int64)
    np.savez_compressed(OUT / (name + ".npz"), **keep)
    print("%-16s N_nms=%d ret=%d -> %d KB" % (name, d["nms_kpts"].shape[1], int(d["out_ret"][0]),
                                               (OUT / (name + ".npz")).stat().st_size // 1024))


def main():
    if not REF.exists():
        sys.exit("build oracle/_ref first: make -C oracle/refbuild")
    OUT.mkdir(parents=True, exist_ok=True)
    for c in CASES:
        run_case(*c)


if __name__ == "__main__":
    main()
