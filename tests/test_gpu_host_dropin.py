"""The C++ drop-in classes (xfeatslam_b200/host: XFextractor, XFBmatcher) called the way the reference's
Frame / Tracking call the originals, compared with the reference's own packed output (golden) and the
C matcher oracle.  Compiled here against the cv stand-in header (OpenCV C++ is not in the image)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import matcher_oracle as mo
from xfeatslam_b200.frames import synthetic_frame, synthetic_pair

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parents[1]
GOLD = REPO / "tests" / "golden"


@pytest.fixture(scope="module")
def driver():
    out = REPO / "tests" / "host" / "_build"
    out.mkdir(exist_ok=True)
    exe = out / "host_dropin_driver"
    host = REPO / "xfeatslam_b200" / "host"
    cmd = ["g++", "-std=c++17", "-O2", "-I", str(REPO / "oracle" / "refbuild" / "shim"), "-I", str(REPO / "include"), "-I", str(host),
           str(REPO / "tests" / "host" / "host_dropin_driver.cc"), str(host / "XFextractor.cc"), str(host / "XFBmatcher.cc"),
           "-L", str(REPO / "xfeatslam_b200" / "lib"), "-lxfeat_b200", "-Wl,-rpath," + str(REPO / "xfeatslam_b200" / "lib"), "-o", str(exe)]
    subprocess.run(cmd, check=True)
    return exe


def canon(kp3, desc):
    valid = kp3[:, 2] > 0
    k, d = kp3[valid], desc[valid]
    order = np.lexsort((k[:, 0], k[:, 1], -k[:, 2].astype(np.float64)))
    return k[order], d[order]


@pytest.mark.parametrize("name", ["vga_top1000", "mono_96x128", "small_64x96"])
def test_xfextractor_operator_matches_reference_packing(driver, tmp_path, name):
    z = np.load(GOLD / (name + ".npz"))
    idx, H, W, nfeat, l0, l1 = [int(v) for v in z["meta"]]
    frame = synthetic_frame(idx, H, W)
    fp = tmp_path / "f.u8"
    frame.tofile(fp)
    pre = tmp_path / "out"
    subprocess.run([str(driver), "extract", str(fp), str(H), str(W), str(nfeat), str(l0), str(l1), str(pre)], check=True)
    kp = np.fromfile(str(pre) + ".kp", np.float32).reshape(-1, 7)
    ds = np.fromfile(str(pre) + ".desc", np.float32).reshape(-1, 64)
    meta = np.fromfile(str(pre) + ".meta", np.float32)
    gk, gd = z["out_keypoints"], z["out_descriptors"]
    assert int(meta[0]) == int(z["out_ret"][0])                       # monoIndex
    assert kp.shape[0] == nfeat and ds.shape == gd.shape               # always nfeatures entries (SURVEY finding)
    assert int(meta[3]) == 8 and abs(meta[4] - 1.2) < 1e-6 and abs(meta[5] - 1.2 ** 7) < 1e-4
    filled = kp[:, 2] > 0
    n_ref = int((gk[:, 2] > 0).sum())
    assert abs(int(filled.sum()) - n_ref) <= max(2, n_ref // 100)
    ret, nf = int(meta[0]), int(filled.sum())
    # monoIndex keypoints from index 0 up, the lapping-area ones (x in [lap0, lap1]) from nfeatures-1 down
    assert filled[:ret].all() and filled[nfeat - (nf - ret):].all() and not filled[ret:nfeat - (nf - ret)].any()
    if l1 >= W:
        assert ret == 0                                                # mono: every keypoint takes the lapping branch
    else:
        assert np.all(kp[nfeat - (nf - ret):, 0] == 0) and np.all(kp[:ret, 0] > 0)   # RGB-D {0,0}: only x == 0 goes to the back
    assert np.all(kp[filled, 3] == 1) and np.all(kp[filled, 4] == -1) and np.all(kp[:, 5] == 0) and np.all(kp[:, 6] == -1)
    assert np.all(kp[~filled, :3] == 0) and np.all(ds[~filled] == 0)   # phantom rows
    ck, cd = canon(kp[:, :3], ds)
    rk, rd = canon(gk, gd)
    got = {(int(x), int(y)): i for i, (x, y) in enumerate(ck[:, :2])}
    pairs = [(got[(int(x), int(y))], j) for j, (x, y) in enumerate(rk[:, :2]) if (int(x), int(y)) in got]
    assert len(pairs) >= 0.97 * n_ref                                  # set differs only at the k-th score boundary
    gi = np.array([p[0] for p in pairs]); ri = np.array([p[1] for p in pairs])
    np.testing.assert_allclose(ck[gi, 2], rk[ri, 2], atol=5e-5, rtol=0)
    np.testing.assert_allclose(cd[gi], rd[ri], atol=1e-4, rtol=0)


def test_search_for_initialization_and_match_mirror(driver, tmp_path):
    from xfeatslam_b200.capi import XFeatB200
    fa, fb = synthetic_pair(3, 480, 640, shift=(9, 4))
    ctx = XFeatB200(max_h=480, max_w=640, max_batch=2, max_topk=1500)
    o = ctx.extract(np.stack([fa, fb]), 1500)
    na, nb = int(o["n_valid"][0]), int(o["n_valid"][1])
    dA, dB, kA, kB = o["desc"][0][:na], o["desc"][1][:nb], o["kpts"][0][:na], o["kpts"][1][:nb]
    paths = {}
    for nm, arr in (("dA", dA), ("kA", kA), ("dB", dB), ("kB", kB)):
        paths[nm] = tmp_path / (nm + ".f32")
        np.ascontiguousarray(arr, np.float32).tofile(paths[nm])
    out_i, out_p = tmp_path / "m.i32", tmp_path / "prev.f32"
    subprocess.run([str(driver), "init", str(paths["dA"]), str(na), str(paths["kA"]), str(paths["dB"]), str(nb), str(paths["kB"]), "640", "480",
                    "100", "0.9", str(out_i), str(out_p)], check=True)
    res = np.fromfile(out_i, np.int32)
    n, m12 = int(res[0]), res[1:1 + na]
    n_want, m_want, prev_want = mo.search_for_initialization(dA, kA, dB, kB, 640, 480, kA.copy(), window=100, ratio=0.9, th_low=100)
    assert n == n_want and np.array_equal(m12, m_want)                 # bit-exact replay of src/ORBmatcher.cc:833-948
    np.testing.assert_array_equal(np.fromfile(out_p, np.float32).reshape(-1, 2), prev_want)
    assert n > 100
    good = m12 >= 0
    shift = kA[good] - kB[m12[good]]
    assert np.all(np.abs(np.median(shift, axis=0) - np.array([9, 4])) <= 1)          # recovers the synthetic translation
    # ORBmatcher::match slot: mutual nearest neighbours
    nm = int(res[1 + na])
    pairs = res[2 + na:2 + na + 2 * nm].reshape(-1, 2)
    bi, bd, sd, ri, rd = mo.bruteforce(dA, dB)
    want = [(i, int(bi[i])) for i in range(na) if bi[i] >= 0 and ri[bi[i]] == i]
    assert [tuple(p) for p in pairs.tolist()] == want
    ctx.close()
