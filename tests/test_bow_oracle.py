"""oracle.mo_bow_transform (C restatement of TemplatedVocabulary::transform + FORB::distance) against a naive Python walk on a
synthetic vocabulary; optionally (XFB_TEST_ORBVOC=1, reference tree mounted) on the reference's real ORBvoc.txt."""
import os
from pathlib import Path

import numpy as np
import pytest

from oracle import matcher_oracle as mo
from tools import orbvoc


def naive(desc, v, levelsup):
    cs, ci, nd, L = v["child_start"], v["child_index"], v["node_desc"], v["L"]
    leaf = np.zeros(len(desc), np.int32); nid = np.zeros(len(desc), np.int32)
    for i, d in enumerate(desc):
        f = np.frombuffer(np.ascontiguousarray(d, np.float32).tobytes()[:32], np.uint8)
        node, level = 0, 0
        nid[i] = 0 if L - levelsup <= 0 else -1
        while cs[node + 1] > cs[node]:
            level += 1
            kids = ci[cs[node]:cs[node + 1]]
            dist = [int(np.unpackbits(f ^ nd[c]).sum()) for c in kids]
            node = int(kids[int(np.argmin(dist))])            # argmin = first minimum, like `if (d < best_d)`
            if level == L - levelsup:
                nid[i] = node
        leaf[i] = node
    return leaf, nid


@pytest.mark.parametrize("k,L,levelsup", [(10, 3, 1), (4, 4, 2), (3, 2, 4)])
def test_bow_transform_matches_naive_walk(k, L, levelsup):
    v = orbvoc.synthetic(k=k, L=L, seed=k + L)
    rng = np.random.RandomState(1)
    desc = rng.randn(200, 64).astype(np.float32)
    desc /= np.linalg.norm(desc, axis=1, keepdims=True)
    desc[7] = 0
    leaf, nid = mo.bow_transform(desc, v["node_desc"], v["child_start"], v["child_index"], v["L"], levelsup)
    wl, wn = naive(desc, v, levelsup)
    assert np.array_equal(leaf, wl) and np.array_equal(nid, wn)
    assert np.all(v["is_leaf"][leaf] == 1)


def test_text_format_round_trip():
    v = orbvoc.synthetic(k=3, L=2, seed=5)
    lines = ["3 2 0 0"]
    parent = np.zeros(v["node_desc"].shape[0], np.int32)
    for p in range(v["node_desc"].shape[0]):
        for c in v["child_index"][v["child_start"][p]:v["child_start"][p + 1]]:
            parent[c] = p
    for i in range(1, v["node_desc"].shape[0]):
        lines.append("%d %d %s %r" % (parent[i], v["is_leaf"][i], " ".join(str(int(b)) for b in v["node_desc"][i]), float(v["weight"][i])))
    w = orbvoc.from_lines(lines)
    for key in ("node_desc", "child_start", "child_index", "word_id", "weight"):
        assert np.array_equal(v[key], w[key]), key


@pytest.mark.skipif(os.environ.get("XFB_TEST_ORBVOC") != "1" or not Path("/root/reference/Vocabulary/ORBvoc.txt.tar.gz").exists(),
                    reason="needs the reference tree and XFB_TEST_ORBVOC=1 (parsing the 145 MB text takes minutes)")
def test_real_orbvoc_shape():
    v = orbvoc.load("/root/reference/Vocabulary/ORBvoc.txt.tar.gz")
    assert v["k"] == 10 and v["L"] == 6
    z = np.load(Path(__file__).resolve().parent / "golden" / "vga_top1000.npz")
    d = z["out_descriptors"][:200]
    leaf, nid = mo.bow_transform(d, v["node_desc"], v["child_start"], v["child_index"], v["L"], 4)
    assert np.all(v["is_leaf"][leaf] == 1) and np.all(nid > 0)
    # the pin on the real vocabulary: word / node ids the reference's own DBoW2 returned for these descriptors (tools/pin_orbvoc.py)
    g = np.load(Path(__file__).resolve().parent / "golden" / "dbow2_orbvoc_vga1000.npz")
    d = np.ascontiguousarray(z["out_descriptors"][:1000], np.float32)
    leaf, nid = mo.bow_transform(d, v["node_desc"], v["child_start"], v["child_index"], v["L"], int(g["meta"][4]))
    assert np.array_equal(v["word_id"][leaf], g["leaf"][:, 0]) and np.array_equal(nid, g["leaf"][:, 1])


# ---- the pin: the reference's OWN DBoW2 (thirdparty/DBoW2 compiled unchanged -> oracle/_ref/ref_dbow2) -------------------------
GOLD = Path(__file__).resolve().parent / "golden"


def assemble(desc, voc, levelsup):
    """BowVector / FeatureVector of TemplatedVocabulary::transform (TemplatedVocabulary.h:1147-1193, TF_IDF + L1) from the oracle's walk."""
    wl, wn = mo.bow_transform(desc, voc["node_desc"], voc["child_start"], voc["child_index"], voc["L"], levelsup)
    v, fv = {}, {}
    for i in range(len(desc)):
        w = float(voc["weight"][wl[i]])
        if w > 0:
            wid = int(voc["word_id"][wl[i]])
            v[wid] = v.get(wid, 0.0) + w
            fv.setdefault(int(wn[i]), []).append(i)
    norm = 0.0
    for k in sorted(v):
        norm += abs(v[k])
    bow = np.array([[k, v[k] / norm] for k in sorted(v)], np.float64).reshape(-1, 2)
    flat = []
    for k in sorted(fv):
        flat += [k, len(fv[k])] + fv[k]
    return wl, wn, bow, np.array(flat, np.int32)


@pytest.mark.parametrize("name", ["dbow2_k10L4", "dbow2_k4L5", "dbow2_k3L2"])
def test_oracle_walk_pinned_against_reference_dbow2_golden(name):
    """tests/golden/dbow2_*.npz = outputs of the reference's own TemplatedVocabulary<FORB>::transform (tools/make_golden_dbow2.py):
    word id, node id, BowVector (bit-identical doubles) and FeatureVector must be reproduced by the oracle restatement."""
    from tools.make_golden_dbow2 import descriptors
    z = np.load(GOLD / (name + ".npz"))
    k, L, seed, n, levelsup = [int(x) for x in z["meta"]]
    voc = orbvoc.synthetic(k=k, L=L, seed=seed)
    desc = descriptors(seed, n)
    wl, wn, bow, fv = assemble(desc, voc, levelsup)
    assert np.array_equal(voc["word_id"][wl], z["leaf"][:, 0])                       # WordId of every feature
    assert np.array_equal(wn, z["leaf"][:, 1])                                       # NodeId at level L - levelsup
    assert np.array_equal((voc["weight"][wl] > 0).astype(np.int32), z["leaf"][:, 2])
    assert np.array_equal(bow, z["bow"])                                             # same doubles, same order
    assert np.array_equal(fv, z["fv"])


@pytest.mark.skipif(not (Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "ref_dbow2").exists(), reason="oracle/_ref/ref_dbow2 is built where the reference tree is mounted")
def test_oracle_walk_against_reference_dbow2_live():
    """The same comparison on fresh inputs, running the reference binary now (other seeds, levelsup = 0 .. L + 1)."""
    from tools.make_golden_dbow2 import descriptors, run_reference
    voc = orbvoc.synthetic(k=5, L=3, seed=11)
    desc = descriptors(77, 300)
    for levelsup in (0, 1, 2, 3, 4):
        leaf, bow_ref, fv_ref = run_reference(voc, desc, levelsup)
        wl, wn, bow, fv = assemble(desc, voc, levelsup)
        assert np.array_equal(voc["word_id"][wl], leaf[:, 0]) and np.array_equal(wn, leaf[:, 1])
        assert np.array_equal(bow, bow_ref) and np.array_equal(fv, fv_ref)
