#!/usr/bin/env python
"""DBoW2 vocabulary text files (the reference's Vocabulary/ORBvoc.txt) -> the flat arrays xfb_vocab_load takes, and a
synthetic vocabulary generator for tests.  Follows TemplatedVocabulary::loadFromTextFile
(thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1420): first line `k L scoring weighting`, then one node per line
`parent isLeaf d0 .. d31 weight`; node ids count up from 1 in file order (0 = root), leaves get word ids in file order."""
import sys
import tarfile

import numpy as np


def from_lines(lines):
    k, L, scoring, weighting = [int(v) for v in lines[0].split()[:4]]
    rows = [ln.split() for ln in lines[1:] if ln.strip()]
    n = len(rows) + 1
    parent = np.zeros(n, np.int32); is_leaf = np.zeros(n, np.uint8); desc = np.zeros((n, 32), np.uint8); weight = np.zeros(n, np.float64)
    for i, r in enumerate(rows, start=1):
        parent[i] = int(r[0]); is_leaf[i] = int(r[1]) > 0
        desc[i] = np.array(r[2:34], dtype=np.int64).astype(np.uint8)
        weight[i] = float(r[34])
    return finish(k, L, scoring, weighting, parent, is_leaf, desc, weight)


def finish(k, L, scoring, weighting, parent, is_leaf, desc, weight):
    n = parent.shape[0]
    # m_nodes[pid].children.push_back(nid) in file order -> CSR with a stable sort by parent
    order = np.argsort(parent[1:], kind="stable").astype(np.int32) + 1
    counts = np.bincount(parent[1:], minlength=n)
    child_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    word_id = np.full(n, -1, np.int32)
    word_id[is_leaf > 0] = np.arange(int((is_leaf > 0).sum()), dtype=np.int32)
    return {"k": k, "L": L, "scoring": scoring, "weighting": weighting, "node_desc": desc, "child_start": child_start, "child_index": order,
            "weight": weight, "word_id": word_id, "is_leaf": is_leaf}


def load(path):
    """path: ORBvoc.txt or ORBvoc.txt.tar.gz"""
    if str(path).endswith(".tar.gz"):
        with tarfile.open(path) as tf:
            f = tf.extractfile(tf.getmembers()[0])
            lines = f.read().decode().split("\n")
    else:
        lines = open(path).read().split("\n")
    return from_lines(lines)


def synthetic(k=10, L=3, seed=0, stop_fraction=0.1):
    """A full k-ary tree of depth L with random 32-byte node words, nodes numbered in the order a file would list them
    (breadth first here), random idf weights with a few stopped (weight 0) words."""
    rng = np.random.RandomState(seed)
    parent = [0]; level = [0]
    frontier = [0]
    for lv in range(1, L + 1):
        nxt = []
        for p in frontier:
            for _ in range(k):
                parent.append(p); level.append(lv); nxt.append(len(parent) - 1)
        frontier = nxt
    parent = np.array(parent, np.int32); level = np.array(level)
    n = parent.shape[0]
    is_leaf = (level == L).astype(np.uint8)
    desc = rng.randint(0, 256, (n, 32)).astype(np.uint8)
    desc[0] = 0                                               # the root has no word (it is not in the file)
    weight = np.where(is_leaf > 0, rng.rand(n) * 5, 0.0)
    weight[(rng.rand(n) < stop_fraction) & (is_leaf > 0)] = 0.0
    return finish(k, L, 0, 0, parent, is_leaf, desc, weight)


if __name__ == "__main__":
    v = load(sys.argv[1])
    print("k=%d L=%d nodes=%d words=%d" % (v["k"], v["L"], v["node_desc"].shape[0], int((v["word_id"] >= 0).sum())))
