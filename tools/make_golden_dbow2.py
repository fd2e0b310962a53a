#!/usr/bin/env python
"""Golden vectors for the vocabulary tree walk, produced by the reference's OWN DBoW2 (oracle/_ref/ref_dbow2 =
thirdparty/DBoW2 compiled unchanged, oracle/refbuild/Makefile): tests/golden/dbow2_k10L4.npz.

  python tools/make_golden_dbow2.py        # needs /root/reference (this container); the fixture travels to the GPU box

Inputs are regenerated from seeds by the tests (tools/orbvoc.synthetic + a seeded descriptor set), so only the reference's
OUTPUTS are stored: per descriptor (word id, node id at level L - levelsup, weight > 0), the BowVector and the FeatureVector."""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from tests.host_cases import vocabulary_text  # noqa: E402
from tools import orbvoc  # noqa: E402

CASES = {"dbow2_k10L4": dict(k=10, L=4, seed=3, n=600, levelsup=2), "dbow2_k4L5": dict(k=4, L=5, seed=8, n=400, levelsup=4),
         "dbow2_k3L2": dict(k=3, L=2, seed=5, n=100, levelsup=4)}


def descriptors(seed, n):
    """Unit rows like XFeat descriptors, plus the phantom all-zero rows the reference keeps."""
    rng = np.random.RandomState(1000 + seed)
    d = rng.randn(n, 64).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[5] = 0; d[n - 1] = 0
    return d


def run_reference(voc, desc, levelsup):
    exe = REPO / "oracle" / "_ref" / "ref_dbow2"
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        (td / "voc.txt").write_text(vocabulary_text(voc))
        desc.tofile(td / "d.f32")
        subprocess.run([str(exe), str(td / "voc.txt"), str(td / "d.f32"), str(len(desc)), str(levelsup), str(td / "o")], check=True, capture_output=True)
        leaf = np.fromfile(str(td / "o.leaf"), np.int32).reshape(-1, 3)
        bow = np.fromfile(str(td / "o.bow"), np.float64).reshape(-1, 2)
        fv = np.fromfile(str(td / "o.fv"), np.int32)
    return leaf, bow, fv


if __name__ == "__main__":
    for name, c in CASES.items():
        voc = orbvoc.synthetic(k=c["k"], L=c["L"], seed=c["seed"])
        leaf, bow, fv = run_reference(voc, descriptors(c["seed"], c["n"]), c["levelsup"])
        np.savez_compressed(REPO / "tests" / "golden" / (name + ".npz"), leaf=leaf, bow=bow, fv=fv,
                            meta=np.array([c["k"], c["L"], c["seed"], c["n"], c["levelsup"]], np.int32))
        print(name, leaf.shape, bow.shape, fv.shape)
