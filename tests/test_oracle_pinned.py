"""The torch-CPU restatement (oracle/xfeat_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, written by tools/make_golden.py from oracle/_ref/ref_xfeat = the reference's
XFeat.cc + XFextractor.cc compiled unchanged).  This is what pins the oracle."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import xfeat_oracle as xo
from xfeatslam_b200.frames import synthetic_frame

GOLD = Path(__file__).parent / "golden"
FULL = ["small_64x96", "resize_100x140", "mono_96x128"]
FINAL = ["vga_top4096", "vga_top1000", "hd720_top1000"]


def load(name):
    z = np.load(GOLD / (name + ".npz"))
    d = {k.replace("__", "."): z[k] for k in z.files}
    idx, H, W, nfeat, l0, l1 = [int(v) for v in d.pop("meta")]
    return d, synthetic_frame(idx, H, W), nfeat, (l0, l1)


def canon(kp3, desc):
    """Canonical order for the reference's unstable argsort: (score desc, y, x) over non-phantom rows."""
    valid = kp3[:, 2] > 0
    k, d = kp3[valid], desc[valid]
    order = np.lexsort((k[:, 0], k[:, 1], -k[:, 2].astype(np.float64)))
    return k[order], d[order]


@pytest.mark.parametrize("name", FULL)
def test_every_intermediate_matches_reference(name, weights):
    gold, frame, nfeat, lap = load(name)
    keep = {}
    xo.detect_and_compute(frame, weights, nfeat, keep)
    checked = 0
    for key, ref in gold.items():
        if key.startswith("out_") or key in ("desc_all",) or key.startswith("skip1"):
            continue
        got = keep[key].numpy() if torch.is_tensor(keep[key]) else np.asarray(keep[key])
        if key == "nms_kpts":
            assert np.array_equal(got, ref.astype(np.int64)), key
        else:
            assert got.shape == ref.shape, key
            # same ATen operators => normally bit-identical; allow thread-count dependent summation order
            np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5, err_msg=key)
        checked += 1
    assert checked >= 25


@pytest.mark.parametrize("name", FULL + FINAL)
def test_final_outputs_match_reference(name, weights):
    gold, frame, nfeat, lap = load(name)
    kp, sc, ds = xo.detect_and_compute(frame, weights, nfeat)
    pk, pd, mono = xo.pack_reference_layout(kp, sc, ds, nfeat, lap)
    assert mono == int(gold["out_ret"][0])
    assert pk.shape == gold["out_keypoints"].shape and pd.shape == gold["out_descriptors"].shape
    gk, gd = canon(gold["out_keypoints"], gold["out_descriptors"])
    ok, od = canon(pk, pd)
    assert np.array_equal(ok[:, :2], gk[:, :2])
    np.testing.assert_allclose(ok[:, 2], gk[:, 2], rtol=0, atol=1e-6)
    np.testing.assert_allclose(od, gd, rtol=0, atol=1e-5)
    # phantom rows (src/XFextractor.cc:310-311): default keypoints / zero descriptors
    n_valid = int((gold["out_keypoints"][:, 2] > 0).sum())
    assert n_valid == len(kp)


def test_reference_quirks_are_restated(weights):
    """SURVEY.md findings: output always has nfeatures rows; mono lapping fills from the back."""
    gold, frame, nfeat, lap = load("mono_96x128")
    assert lap == (0, 1000) and int(gold["out_ret"][0]) == 0          # every keypoint took the 'stereo' branch
    k = gold["out_keypoints"]
    assert k.shape[0] == nfeat
    nz = np.nonzero(k[:, 2] > 0)[0]
    assert nz.max() == nfeat - 1                                      # filled from the back
    assert k[nz.max(), 2] >= k[nz.min(), 2]                           # best score last
    # last row / column keypoints never survive (nearest sampling falls off the map)
    g2, f2, n2, _ = load("vga_top4096")
    kk = g2["out_keypoints"]
    assert not np.any((kk[:, 0] == 639) | (kk[:, 1] == 479))


def test_live_reference_binary_if_present(weights, tmp_path):
    """When oracle/_ref was built (this container), also compare against a fresh run of it."""
    import subprocess
    from oracle.refdump import read_dump
    ref = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "ref_xfeat"
    if not ref.exists():
        pytest.skip("oracle/_ref not built")
    frame = synthetic_frame(11, 96, 96)
    fp, op = tmp_path / "f.u8", tmp_path / "o.bin"
    frame.tofile(fp)
    subprocess.run([str(ref), "dump", str(fp), "96", "96", "200", "0", "0", str(op), "2"], check=True, capture_output=True)
    d = read_dump(op)
    keep = {}
    xo.detect_and_compute(frame, weights, 200, keep)
    for key in ("xn", "feats", "K1", "H1", "K1h", "scores_all"):
        np.testing.assert_allclose(keep[key].numpy(), d[key], rtol=0, atol=2e-5, err_msg=key)
    assert np.array_equal(keep["nms_kpts"].numpy(), d["nms_kpts"])
