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


def _bundle(path, arrays):
    """[name 16 bytes][dtype f / i / b][pad 7][count int64][data] per array (read by tests/host/host_dropin_driver.cc)."""
    import struct
    with open(path, "wb") as f:
        for name, arr in arrays.items():
            arr = np.asarray(arr)
            if arr.dtype == np.float32:
                code = b"f"
            elif arr.dtype == np.int32:
                code = b"i"
            else:
                arr = arr.astype(np.uint8); code = b"b"
            f.write(name.encode().ljust(16, b"\0") + code + b"\0" * 7 + struct.pack("<q", arr.size))
            f.write(np.ascontiguousarray(arr).tobytes())


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
    ctx.close()
    # the synthetic motion between the two frames, measured (its sign convention is synthetic_pair's business)
    bi, bd, _, _, _ = mo.bruteforce(dA, dB)
    ok = bd < 60
    dx, dy = [float(v) for v in np.median(kB[bi[ok]] - kA[ok], axis=0)]
    assert ok.sum() > 200 and abs(abs(dx) - 7) <= 1 and abs(abs(dy) - 3) <= 1
    # vocabulary nodes: a 10 x 10 grid of "level-2 nodes" (k = 10, L = 6, levelsup = 4 -> 100 nodes) by image position
    nodeA = ((kA[:, 0] // 64).astype(np.int32) * 10 + (kA[:, 1] // 48).astype(np.int32)).astype(np.int32)
    nodeB = (((kB[:, 0] - dx) // 64).astype(np.int32) * 10 + ((kB[:, 1] - dy) // 48).astype(np.int32)).astype(np.int32)
    nodeB = np.clip(nodeB, 0, 99).astype(np.int32)
    nodeA[rng.rand(na) < 0.05] = -1; nodeB[rng.rand(nb) < 0.05] = -1
    goodA = (rng.rand(na) < 0.8); goodB = (rng.rand(nb) < 0.8)
    hasA = (rng.rand(na) < 0.3); hasB = (rng.rand(nb) < 0.3)
    stA = (rng.rand(na) < 0.5); stB = (rng.rand(nb) < 0.5)
    F12 = np.array([[0, 0, 0], [0, 0, -1], [0, 1, -dy]], np.float32)      # x-translation: y2 = y1 + dy on the epipolar line
    ep = np.array([900.0, 240.0], np.float32)
    # projection search: map points = frame-A descriptors projected near their frame-B positions
    nM = 600
    src = rng.randint(0, na, nM)
    proj = (kA[src] + np.array([dx, dy], np.float32) + rng.randn(nM, 2).astype(np.float32)).astype(np.float32)
    level = rng.choice([0, 0, 0, 1, 2], nM).astype(np.int32)
    viewcos = rng.choice([0.9, 0.9995], nM).astype(np.float32)
    in_view = rng.rand(nM) < 0.9; mp_obs = rng.rand(nM) < 0.9
    occupied = rng.rand(nb) < 0.1
    uright = np.where(rng.rand(nb) < 0.5, kB[:, 0] - 25.0, -1.0).astype(np.float32)
    projxr = (proj[:, 0] - 25.0 + 2 * rng.randn(nM)).astype(np.float32)
    sizes = [1, 2, 3, 5, 8, 13, 40]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    dS = dA[rng.randint(0, na, offsets[-1])] + 0.05 * rng.randn(offsets[-1], 64).astype(np.float32)
    dS = (dS / np.linalg.norm(dS, axis=1, keepdims=True)).astype(np.float32)
    arrays = dict(dA=dA, dB=dB, kA=kA, kB=kB, nodeA=nodeA, nodeB=nodeB, goodA=goodA, goodB=goodB, hasmpA=hasA, hasmpB=hasB, stereoA=stA, stereoB=stB,
                  F12=F12.reshape(-1), ep=ep, ratio_kf_f=np.float32([0.7]), ratio_kf_kf=np.float32([0.9]), ratio_proj=np.float32([0.8]),
                  th_proj=np.float32([3.0]), dM=dA[src], dF=dB, kF=kB, proj=proj.reshape(-1), projxr=projxr, viewcos=viewcos, level=level,
                  in_view=in_view, mp_obs=mp_obs, occupied=occupied, uright=uright, img_wh=np.float32([640, 480]), dS=dS, offsets=offsets)
    _bundle(tmp_path / "in.bin", arrays)
    subprocess.run([str(driver), "searches", str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], check=True)
    res = np.fromfile(tmp_path / "out.bin", np.int32)
    pos = 0

    def take():
        nonlocal pos
        n, ln = int(res[pos]), int(res[pos + 1])
        v = res[pos + 2:pos + 2 + ln].copy(); pos += 2 + ln
        return n, v
    n, m = take(); wn, wm = mo.search_by_bow_kf_f(dA, nodeA, goodA, dB, nodeB, ratio=0.7, th_low=100)
    assert n == wn and np.array_equal(m, wm) and n > 50
    n, m = take(); wn, wm = mo.search_by_bow_kf_kf(dA, nodeA, goodA, dB, nodeB, goodB, ratio=0.9, th_low=100)
    assert n == wn and np.array_equal(m, wm) and n > 50
    for coarse in (False, True):
        n, m = take(); wn, wm = mo.search_for_triangulation(dA, nodeA, hasA, stA, kA, dB, nodeB, hasB, stB, kB, F12, ep, only_stereo=False, coarse=coarse)
        assert n == wn and np.array_equal(m, wm) and n > 20
    n, m = take(); wn, wm = mo.search_by_projection(dA[src], in_view, proj, projxr, level, viewcos, mp_obs, dB, kB, occupied, uright, 640, 480, th=3.0,
                                                     scale_factor=1.2, ratio=0.8, th_high=1000)
    assert n == wn and np.array_equal(m, wm) and n > 50
    _, best = take()
    assert np.array_equal(best, mo.distinctive_descriptors(dS, offsets))


def test_vocabulary_transform_on_device(driver, tmp_path):
    """xfb_vocab_load + xfb_bow_transform (csrc/bow.cu) against the C restatement of TemplatedVocabulary::transform / FORB::distance,
    and XFBvocabulary (text loader + BowVector / FeatureVector bookkeeping) against the same assembled in Python."""
    from tools import orbvoc
    from xfeatslam_b200.capi import XFeatB200
    voc = orbvoc.synthetic(k=10, L=4, seed=3)                       # 11 111 nodes, 10 000 words
    frame = synthetic_frame(11, 480, 640)
    ctx = XFeatB200(max_h=480, max_w=640, max_batch=2, max_topk=2000)
    o = ctx.extract(np.stack([frame, synthetic_frame(12, 480, 640)]), 2000)
    n = int(o["n_valid"][0])
    desc = np.ascontiguousarray(o["desc"][0][:n])
    ctx.vocab_load(voc["node_desc"], voc["child_start"], voc["child_index"], voc["L"])
    for levelsup in (4, 2, 0):
        leaf, nid = ctx.bow_transform(desc, levelsup)
        wl, wn = mo.bow_transform(desc, voc["node_desc"], voc["child_start"], voc["child_index"], voc["L"], levelsup)
        assert np.array_equal(leaf, wl) and np.array_equal(nid, wn)
    # every frame of the batch in one launch, descriptors resident on the device
    import torch
    d_leaf = torch.zeros(2, 2000, dtype=torch.int32, device="cuda"); d_nid = torch.zeros(2, 2000, dtype=torch.int32, device="cuda")
    ctx.extract(np.stack([frame, synthetic_frame(12, 480, 640)]), 2000)
    ctx.bow_transform_frames(2, d_leaf.data_ptr(), d_nid.data_ptr())
    torch.cuda.synchronize()
    for b in range(2):
        nb = int(o["n_valid"][b])
        wl, wn = mo.bow_transform(o["desc"][b][:nb], voc["node_desc"], voc["child_start"], voc["child_index"], voc["L"], 2)
        assert np.array_equal(d_leaf[b, :nb].cpu().numpy(), wl) and np.array_equal(d_nid[b, :nb].cpu().numpy(), wn)
        assert np.all(d_leaf[b, nb:].cpu().numpy() == -1)
    ctx.close()
    # the C++ class: text file in, maps out
    lines = ["%d %d 0 0" % (voc["k"], voc["L"])]
    parent = np.zeros(voc["node_desc"].shape[0], np.int32)
    for p in range(voc["node_desc"].shape[0]):
        parent[voc["child_index"][voc["child_start"][p]:voc["child_start"][p + 1]]] = p
    for i in range(1, voc["node_desc"].shape[0]):
        lines.append("%d %d %s %r" % (parent[i], voc["is_leaf"][i], " ".join(str(int(b)) for b in voc["node_desc"][i]), float(voc["weight"][i])))
    (tmp_path / "voc.txt").write_text("\n".join(lines) + "\n")
    desc.tofile(tmp_path / "d.f32")
    subprocess.run([str(driver), "bow", str(tmp_path / "voc.txt"), str(tmp_path / "d.f32"), str(n), "2", str(tmp_path / "out")], check=True)
    meta = np.fromfile(str(tmp_path / "out") + ".meta", np.int32)
    assert list(meta) == [10, 4, 10000]
    bow = np.fromfile(str(tmp_path / "out") + ".bow", np.float64).reshape(-1, 2)
    fv = np.fromfile(str(tmp_path / "out") + ".fv", np.int32)
    # TemplatedVocabulary::transform (TemplatedVocabulary.h:1147-1193) for TF_IDF weighting (0) + L1 scoring (0), from the oracle's walk
    wl, wn = mo.bow_transform(desc, voc["node_desc"], voc["child_start"], voc["child_index"], voc["L"], 2)
    want_v, want_fv = {}, {}
    for i in range(n):
        w = float(voc["weight"][wl[i]])
        if w > 0:
            wid = int(voc["word_id"][wl[i]])
            want_v[wid] = want_v.get(wid, 0.0) + w
            want_fv.setdefault(int(wn[i]), []).append(i)
    norm = sum(abs(x) for x in (want_v[k] for k in sorted(want_v)))
    assert [int(k) for k in bow[:, 0]] == sorted(want_v)
    np.testing.assert_array_equal(bow[:, 1], np.array([want_v[k] / norm for k in sorted(want_v)]))   # same op order -> same doubles
    flat = []
    for k in sorted(want_fv):
        flat += [k, len(want_fv[k])] + want_fv[k]
    assert fv.tolist() == flat and len(want_fv) > 50
