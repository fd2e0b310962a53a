"""Shared cases for the host drop-in classes (xfeatslam_b200/host): the same inputs / oracle expectations are used by the GPU
tests (tests/test_gpu_host_dropin.py: real libxfeat_b200.so) and by the CPU host-logic tests (tests/test_host_logic_cpu.py: the
driver linked against tests/host/cpu_stub.cc)."""
import struct
import subprocess

import numpy as np

from oracle import matcher_oracle as mo


def bundle(path, arrays):
    """[name 16 bytes][dtype f / i / b][pad 7][count int64][data] per array (read by tests/host/host_dropin_driver.cc)."""
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


def run_searches_case(driver, tmp_path, dA, kA, dB, kB, seed=21, frame_to_frame=False):
    """XFBmatcher::SearchByBoW (both overloads), SearchForTriangulation, SearchByProjection and ComputeDistinctiveDescriptors
    against the C restatements of src/ORBmatcher.cc:408-610, :950-1090, :1092-1331, :42-212 and src/MapPoint.cc:329-403."""
    rng = np.random.RandomState(seed)
    na, nb = len(dA), len(dB)
    # the synthetic motion between the two frames, measured (its sign convention is the generator's business)
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
    if frame_to_frame:
        # SearchByProjection(CurrentFrame, LastFrame): the map points of frame A projected into frame B with a slightly wrong pose
        lf_uv = (kA + np.array([dx, dy], np.float32) + 2.0 * rng.randn(na, 2).astype(np.float32)).astype(np.float32)
        lf_valid = (rng.rand(na) < 0.8) & (lf_uv[:, 0] >= 0) & (lf_uv[:, 0] <= 640) & (lf_uv[:, 1] >= 0) & (lf_uv[:, 1] <= 480)
        lf_invzc = (1.0 / (0.5 + 4.0 * rng.rand(na))).astype(np.float32)
        lf_octave = rng.choice([0, 0, 0, 0, 1], na).astype(np.int32)
        lf_obs = rng.rand(na) < 0.9
        lf_modes = np.array([0, 0, 1, 0, 0, 1], np.int32)
        arrays.update(lf_uv=lf_uv.reshape(-1), lf_valid=lf_valid, lf_invzc=lf_invzc, lf_octave=lf_octave, lf_obs=lf_obs, lf_modes=lf_modes,
                      lf_th=np.float32([15.0]), lf_mbf=np.float32([40.0]))
    if frame_to_frame:
        # loop-closing SearchByProjection (Sim3) and Fuse: the same projected map points, windows of 4 / 7 px
        wq_valid = rng.rand(nM) < 0.85
        wq_level = rng.choice([0, 0, 1, 2], nM).astype(np.int32)
        wq_radius = (np.float32(5.0) * np.float32(1.2) ** wq_level).astype(np.float32)
        wq_ur = projxr
        arrays.update(wq_uv=proj.reshape(-1), wq_ur=wq_ur, wq_radius=wq_radius, wq_level=wq_level, wq_valid=wq_valid, wq_ratio=np.float32([1.0]),
                      wq_invsigma2=np.float32([1.0]))
    if frame_to_frame:
        # SearchBySim3: every feature's map point projected into the other keyframe with a slightly wrong Sim3
        s3_uv1 = (kA + np.array([dx, dy], np.float32) + 1.5 * rng.randn(na, 2).astype(np.float32)).astype(np.float32)
        s3_uv2 = (kB - np.array([dx, dy], np.float32) + 1.5 * rng.randn(nb, 2).astype(np.float32)).astype(np.float32)
        s3_ok1 = rng.rand(na) < 0.8; s3_ok2 = rng.rand(nb) < 0.8
        s3_lvl1 = rng.choice([0, 0, 1, 2], na).astype(np.int32); s3_lvl2 = rng.choice([0, 0, 1, 2], nb).astype(np.int32)
        s3_rad1 = (np.float32(7.5) * np.float32(1.2) ** s3_lvl1).astype(np.float32); s3_rad2 = (np.float32(7.5) * np.float32(1.2) ** s3_lvl2).astype(np.float32)
        arrays.update(s3_uv1=s3_uv1.reshape(-1), s3_uv2=s3_uv2.reshape(-1), s3_ok1=s3_ok1, s3_ok2=s3_ok2, s3_lvl1=s3_lvl1, s3_lvl2=s3_lvl2,
                      s3_rad1=s3_rad1, s3_rad2=s3_rad2)
    bundle(tmp_path / "in.bin", arrays)
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
    if frame_to_frame:
        for fwd, bwd in ((0, 0), (1, 0), (0, 1)):
            n, m = take()
            wn, wm = mo.search_by_projection_frames(dA, lf_valid, lf_uv, lf_invzc, lf_octave, lf_obs, dB, kB, occupied, uright, 640, 480, th=15.0,
                                                    scale_factor=1.2, mbf=40.0, forward=bool(fwd), backward=bool(bwd), th_high=1000)
            assert n == wn and np.array_equal(m, wm) and n > 100
        n, m = take()
        wn, wm = mo.search_by_projection_sim3(dA[src], wq_valid, proj, wq_radius, wq_level, dB, kB, occupied, 640, 480, th_low=100, ratio_hamming=1.0)
        assert n == wn and np.array_equal(m, wm) and n > 50
        n, m = take()
        wn, wm = mo.search_by_projection_reloc(dA[src], wq_valid, proj, wq_radius, wq_level, dB, kB, occupied, 640, 480, orb_dist=64)
        assert n == wn and np.array_equal(m, wm) and n > 50
        _, bi_ = take(); _, bd_ = take()
        wbi, wbd = mo.fuse_search(dA[src], wq_valid, proj, wq_ur, wq_radius, wq_level, dB, kB, uright, 640, 480, inv_sigma2_0=1.0)
        assert np.array_equal(bi_, wbi) and np.array_equal(bd_, wbd) and (wbd <= 100).sum() > 50
        n, m = take()
        wn, wm = mo.search_by_sim3(dA, s3_ok1, s3_uv1, s3_rad1, s3_lvl1, dB, kB, dB, s3_ok2, s3_uv2, s3_rad2, s3_lvl2, dA, kA, 640, 480, th_high=1000)
        assert n == wn and np.array_equal(m, wm) and n > 50
    _, best = take()
    assert np.array_equal(best, mo.distinctive_descriptors(dS, offsets))


def run_init_case(driver, tmp_path, dA, kA, dB, kB):
    """XFBmatcher::SearchForInitialization + XFBmatcher::match against src/ORBmatcher.cc:833-948 restated in C and the mutual-NN spec."""
    na, nb = len(dA), len(dB)
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
    # ORBmatcher::match slot: mutual nearest neighbours
    nm = int(res[1 + na])
    pairs = res[2 + na:2 + na + 2 * nm].reshape(-1, 2)
    bi, bd, sd, ri, rd = mo.bruteforce(dA, dB)
    want = [(i, int(bi[i])) for i in range(na) if bi[i] >= 0 and ri[bi[i]] == i]
    assert [tuple(p) for p in pairs.tolist()] == want
    return m12


def vocabulary_text(voc):
    """The text file TemplatedVocabulary::loadFromTextFile reads, for a tools/orbvoc.py vocabulary dict."""
    lines = ["%d %d 0 0" % (voc["k"], voc["L"])]
    parent = np.zeros(voc["node_desc"].shape[0], np.int32)
    for p in range(voc["node_desc"].shape[0]):
        parent[voc["child_index"][voc["child_start"][p]:voc["child_start"][p + 1]]] = p
    for i in range(1, voc["node_desc"].shape[0]):
        lines.append("%d %d %s %r" % (parent[i], voc["is_leaf"][i], " ".join(str(int(b)) for b in voc["node_desc"][i]), float(voc["weight"][i])))
    return "\n".join(lines) + "\n"


def run_bow_case(driver, tmp_path, voc, desc, levelsup=2):
    """XFBvocabulary (text loader + BowVector / FeatureVector bookkeeping) against TemplatedVocabulary::transform
    (TemplatedVocabulary.h:1147-1193, TF_IDF weighting + L1 scoring) assembled in Python from the oracle's tree walk."""
    n = len(desc)
    (tmp_path / "voc.txt").write_text(vocabulary_text(voc))
    np.ascontiguousarray(desc, np.float32).tofile(tmp_path / "d.f32")
    subprocess.run([str(driver), "bow", str(tmp_path / "voc.txt"), str(tmp_path / "d.f32"), str(n), str(levelsup), str(tmp_path / "out")], check=True)
    meta = np.fromfile(str(tmp_path / "out") + ".meta", np.int32)
    assert list(meta) == [voc["k"], voc["L"], int((voc["word_id"] >= 0).sum())]
    bow = np.fromfile(str(tmp_path / "out") + ".bow", np.float64).reshape(-1, 2)
    fv = np.fromfile(str(tmp_path / "out") + ".fv", np.int32)
    wl, wn = mo.bow_transform(desc, voc["node_desc"], voc["child_start"], voc["child_index"], voc["L"], levelsup)
    want_v, want_fv = {}, {}
    for i in range(n):
        w = float(voc["weight"][wl[i]])
        if w > 0:
            wid = int(voc["word_id"][wl[i]])
            want_v[wid] = want_v.get(wid, 0.0) + w
            want_fv.setdefault(int(wn[i]), []).append(i)
    norm = 0.0
    for k in sorted(want_v):                       # plain left-to-right double sum like BowVector::normalize (Python >= 3.12's sum() compensates)
        norm += abs(want_v[k])
    assert [int(k) for k in bow[:, 0]] == sorted(want_v)
    np.testing.assert_array_equal(bow[:, 1], np.array([want_v[k] / norm for k in sorted(want_v)]))   # same op order -> same doubles
    flat = []
    for k in sorted(want_fv):
        flat += [k, len(want_fv[k])] + want_fv[k]
    assert fv.tolist() == flat and len(want_fv) > 50
