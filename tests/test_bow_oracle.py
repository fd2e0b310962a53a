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
