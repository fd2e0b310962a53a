"""HOST logic of the drop-in classes without a GPU: the C++ driver is linked against tests/host/cpu_stub.cc (every distance / tree
walk answered by the C oracle) instead of libxfeat_b200.so, so these tests check visiting order, pair-list construction, the
accept / reject replays and the BowVector / FeatureVector bookkeeping of xfeatslam_b200/host -- the same cases the GPU tests run
against the real library (tests/test_gpu_host_dropin.py)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests import host_cases
from tools import orbvoc

REPO = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def driver_cpu():
    out = REPO / "tests" / "host" / "_build"
    out.mkdir(exist_ok=True)
    exe = out / "host_dropin_driver_cpu"
    host = REPO / "xfeatslam_b200" / "host"
    obj = out / "matcher_oracle.o"
    subprocess.run(["gcc", "-O2", "-std=c99", "-ffp-contract=off", "-c", str(REPO / "oracle" / "matcher_oracle.c"), "-o", str(obj)], check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-DXFB_CPU_STUB", "-I", str(REPO / "oracle" / "refbuild" / "shim"), "-I", str(REPO / "include"),
                    "-I", str(host), str(REPO / "tests" / "host" / "host_dropin_driver.cc"), str(REPO / "tests" / "host" / "cpu_stub.cc"),
                    str(host / "XFBmatcher.cc"), str(host / "XFBvocabulary.cc"), str(obj), "-o", str(exe), "-lm"], check=True)
    return exe


@pytest.fixture(scope="module")
def frame_pair():
    """Frame A = the reference's own keypoints / descriptors of the VGA golden frame; frame B = the same scene moved by (+7, -3)
    with perturbed descriptors, shuffled, plus unrelated keypoints."""
    z = np.load(REPO / "tests" / "golden" / "vga_top4096.npz")
    k, d = z["out_keypoints"], z["out_descriptors"]
    valid = k[:, 2] > 0
    kA, dA = np.ascontiguousarray(k[valid][:1000, :2], np.float32), np.ascontiguousarray(d[valid][:1000], np.float32)
    rng = np.random.RandomState(4)
    keep = rng.permutation(len(kA))[:850]
    dB = dA[keep] + 0.04 * rng.randn(len(keep), 64).astype(np.float32)
    kB = kA[keep] + np.array([7, -3], np.float32)
    extra = rng.randn(150, 64).astype(np.float32)
    dB = np.concatenate([dB, extra])
    dB = (dB / np.linalg.norm(dB, axis=1, keepdims=True)).astype(np.float32)
    kB = np.concatenate([kB, np.stack([rng.randint(0, 640, 150), rng.randint(0, 480, 150)], 1).astype(np.float32)])
    inside = (kB[:, 0] >= 0) & (kB[:, 0] < 640) & (kB[:, 1] >= 0) & (kB[:, 1] < 480)
    return dA, kA, np.ascontiguousarray(dB[inside]), np.ascontiguousarray(kB[inside])


def test_search_for_initialization_and_match(driver_cpu, tmp_path, frame_pair):
    dA, kA, dB, kB = frame_pair
    m12 = host_cases.run_init_case(driver_cpu, tmp_path, dA, kA, dB, kB)
    good = m12 >= 0
    assert np.all(np.median(kB[m12[good]] - kA[good], axis=0) == np.array([7, -3]))


def test_node_gated_and_window_searches(driver_cpu, tmp_path, frame_pair):
    dA, kA, dB, kB = frame_pair
    host_cases.run_searches_case(driver_cpu, tmp_path, dA, kA, dB, kB, frame_to_frame=True)


def test_vocabulary_text_loader_and_transform(driver_cpu, tmp_path, frame_pair):
    voc = orbvoc.synthetic(k=10, L=3, seed=8)
    host_cases.run_bow_case(driver_cpu, tmp_path, voc, frame_pair[0], levelsup=1)
