#!/usr/bin/env python
"""One-off pin on the reference's REAL vocabulary (Vocabulary/ORBvoc.txt.tar.gz, k = 10, L = 6, 1 082 073 nodes): the reference's own
DBoW2 (oracle/_ref/ref_dbow2) against the oracle walk (oracle mo_bow_transform) on 1000 XFeat descriptors of the VGA golden frame,
levelsup = 4 as in Frame::ComputeBoW (src/Frame.cc:936).  Writes tests/golden/dbow2_orbvoc_vga1000.npz (the reference's word / node ids).

  tar xzf /root/reference/Vocabulary/ORBvoc.txt.tar.gz -C /tmp/orbvoc && python tools/pin_orbvoc.py /tmp/orbvoc/ORBvoc.txt"""
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from oracle import matcher_oracle as mo  # noqa: E402
from tools import orbvoc  # noqa: E402

if __name__ == "__main__":
    voc_txt = sys.argv[1]
    tmp = Path(voc_txt).parent
    z = np.load(REPO / "tests" / "golden" / "vga_top1000.npz")
    d = np.ascontiguousarray(z["out_descriptors"][:1000], np.float32)
    d.tofile(tmp / "d.f32")
    t = time.time()
    r = subprocess.run([str(REPO / "oracle" / "_ref" / "ref_dbow2"), voc_txt, str(tmp / "d.f32"), "1000", "4", str(tmp / "o")], capture_output=True, text=True, check=True)
    print(r.stdout.strip(), "reference DBoW2: %.1f s" % (time.time() - t))
    leaf = np.fromfile(str(tmp / "o.leaf"), np.int32).reshape(-1, 3)
    t = time.time()
    v = orbvoc.load(voc_txt)
    print("parsed in %.0f s: k=%d L=%d nodes=%d" % (time.time() - t, v["k"], v["L"], v["node_desc"].shape[0]))
    wl, wn = mo.bow_transform(d, v["node_desc"], v["child_start"], v["child_index"], v["L"], 4)
    ok_w, ok_n = np.array_equal(v["word_id"][wl], leaf[:, 0]), np.array_equal(wn, leaf[:, 1])
    print("word ids equal:", ok_w, "| node ids equal:", ok_n, "| distinct level-2 nodes:", len(set(wn.tolist())))
    assert ok_w and ok_n
    np.savez_compressed(REPO / "tests" / "golden" / "dbow2_orbvoc_vga1000.npz", leaf=leaf, meta=np.array([10, 6, 0, 1000, 4], np.int32))
