"""The C++ drop-in classes (xfeatslam_b200/host: XFextractor, XFBmatcher) called the way the reference's
Frame / Tracking call the originals, compared with the reference's own packed output (golden) and the
C matcher oracle.  Compiled here against the cv stand-in header (OpenCV C++ is not in the image)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import matcher_oracle as mo
from tests import host_cases
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
           str(REPO / "tests" / "host" / "host_dropin_driver.cc"), str(host / "XFextractor.cc"), str(host / "XFBmatcher.cc"), str(host / "XFBvocabulary.cc"),
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
    ctx.close()
    m12 = host_cases.run_init_case(driver, tmp_path, dA, kA, dB, kB)
    good = m12 >= 0
    shift = kA[good] - kB[m12[good]]
    assert np.all(np.abs(np.median(shift, axis=0) - np.array([9, 4])) <= 1)          # recovers the synthetic translation


def test_node_gated_and_window_searches_replay_the_oracle(driver, tmp_path):
    """XFBmatcher::SearchByBoW (both overloads), SearchForTriangulation, SearchByProjection and ComputeDistinctiveDescriptors
    (host loops over ONE xfb_distance_pairs launch each) against the C restatements of src/ORBmatcher.cc:408-610, :950-1090,
    :1092-1331, :42-212 and src/MapPoint.cc:329-403 -- bit-exact."""
    from xfeatslam_b200.capi import XFeatB200
    rng = np.random.RandomState(21)
    fa, fb = synthetic_pair(5, 480, 640, shift=(7, -3))
    ctx = XFeatB200(max_h=480, max_w=640, max_batch=2, max_topk=1000)
    o = ctx.extract(np.stack([fa, fb]), 1000)
    na, nb = int(o["n_valid"][0]), int(o["n_valid"][1])
    dA, dB, kA, kB = o["desc"][0][:na].copy(), o["desc"][1][:nb].copy(), o["kpts"][0][:na].copy(), o["kpts"][1][:nb].copy()
    # sparse pair list primitive itself
    ia = rng.randint(0, na, 5000).astype(np.int32); ib = rng.randint(0, nb, 5000).astype(np.int32)
    want = np.array([mo.descriptor_distance(dA[i], dB[j]) for i, j in zip(ia[:500], ib[:500])], np.int32)
    got = ctx.distance_pairs(dA, dB, ia, ib)
    assert np.array_equal(got[:500], want)
    assert np.array_equal(got, mo.distance_matrix(dA, dB)[ia, ib])
    assert len(ctx.distance_pairs(dA, dB, ia[:0], ib[:0])) == 0                       # empty list
    with pytest.raises(RuntimeError):
        ctx.distance_pairs(dA, dB, np.array([na], np.int32), np.array([0], np.int32))  # index out of range -> XFB_ERR_ARG
    ctx.close()
    # frame_to_frame=True: every M5 replay (SearchByProjection x5, Fuse x2, SearchBySim3) runs over the B200's distances too
    host_cases.run_searches_case(driver, tmp_path, dA, kA, dB, kB, frame_to_frame=True)


def test_vocabulary_transform_on_device(driver, tmp_path):
    """xfb_vocab_load + xfb_bow_transform (csrc/bow.cu) against the C restatement of TemplatedVocabulary::transform / FORB::distance,
    and XFBvocabulary (text loader + BowVector / FeatureVector bookkeeping) against the same assembled in Python."""
    import torch
    from tools import orbvoc
    from xfeatslam_b200.capi import XFeatB200
    voc = orbvoc.synthetic(k=10, L=4, seed=3)                       # 11 111 nodes, 10 000 words
    frames = np.stack([synthetic_frame(11, 480, 640), synthetic_frame(12, 480, 640)])
    ctx = XFeatB200(max_h=480, max_w=640, max_batch=2, max_topk=2000)
    o = ctx.extract(frames, 2000)
    n = int(o["n_valid"][0])
    desc = np.ascontiguousarray(o["desc"][0][:n])
    ctx.vocab_load(voc["node_desc"], voc["child_start"], voc["child_index"], voc["L"])
    for levelsup in (4, 2, 0):
        leaf, nid = ctx.bow_transform(desc, levelsup)
        wl, wn = mo.bow_transform(desc, voc["node_desc"], voc["child_start"], voc["child_index"], voc["L"], levelsup)
        assert np.array_equal(leaf, wl) and np.array_equal(nid, wn)
    # every frame of the batch in one launch, descriptors resident on the device
    d_leaf = torch.zeros(2, 2000, dtype=torch.int32, device="cuda"); d_nid = torch.zeros(2, 2000, dtype=torch.int32, device="cuda")
    ctx.extract(frames, 2000)
    ctx.bow_transform_frames(2, d_leaf.data_ptr(), d_nid.data_ptr())
    torch.cuda.synchronize()
    for b in range(2):
        nb = int(o["n_valid"][b])
        wl, wn = mo.bow_transform(o["desc"][b][:nb], voc["node_desc"], voc["child_start"], voc["child_index"], voc["L"], 2)
        assert np.array_equal(d_leaf[b, :nb].cpu().numpy(), wl) and np.array_equal(d_nid[b, :nb].cpu().numpy(), wn)
        assert np.all(d_leaf[b, nb:].cpu().numpy() == -1)
    ctx.close()
    host_cases.run_bow_case(driver, tmp_path, voc, desc, levelsup=2)
