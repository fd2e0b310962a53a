"""oracle/geom_oracle.py pinned against python-opencv's cv2.undistortPoints (the third-party routine behind
Frame::UndistortKeyPoints / ComputeImageBounds), and the host half of the shared undistort code (xfb_image_bounds needs no GPU)."""
import numpy as np
import pytest

from oracle import geom_oracle as go

cv2 = pytest.importorskip("cv2")
TUM1 = dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, k1=0.262383, k2=-0.953104, p1=-0.005358, p2=0.002628, k3=1.163314)


def cv_undistort(xy, cam):
    K = np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1]], np.float32)
    dist = np.array(cam[4:9], np.float32).reshape(5, 1)
    return cv2.undistortPoints(np.asarray(xy, np.float32).reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_undistort_is_bit_identical_to_opencv(seed):
    rng = np.random.RandomState(seed)
    if seed == 0:
        cam = go.camera(bf=40.0, **TUM1)                           # examples/RGB-D/TUM1.yaml
    else:
        cam = go.camera(400 + 300 * rng.rand(), 400 + 300 * rng.rand(), 300 + 40 * rng.rand(), 220 + 40 * rng.rand(), 0.3 * rng.randn(),
                        0.3 * rng.randn(), 0.01 * rng.randn(), 0.01 * rng.randn(), 0.2 * rng.randn())
    xy = np.stack([rng.randint(0, 640, 5000), rng.randint(0, 480, 5000)], 1).astype(np.float32)
    xy[:4] = [[0, 0], [640, 0], [0, 480], [640, 480]]
    got, want = go.undistort(xy, cam), cv_undistort(xy, cam)
    assert np.array_equal(got, want), float(np.abs(got - want).max())


def test_zero_distortion_copies_keypoints():
    cam = go.camera(500, 500, 320, 240)
    xy = np.float32([[3, 4], [100.5, 7]])
    assert np.array_equal(go.undistort(xy, cam), xy)
    assert list(go.image_bounds(cam, 640, 480)[10:14]) == [0, 0, 640, 480]


def test_host_image_bounds_matches_oracle():
    """xfb_image_bounds runs the SAME undistort routine as the kernel, compiled for the host: no GPU needed."""
    from xfeatslam_b200 import capi
    for cam in (go.camera(bf=40.0, **TUM1), go.camera(535.4, 539.2, 320.1, 247.6, 0.0, 0, 0, 0, 0), go.camera(520.9, 521.0, 325.1, 249.7, 0.2312, -0.7849, -0.0033, -0.0001, 0.9172)):
        got = capi.image_bounds(cam.copy(), 640, 480)
        want = go.image_bounds(cam, 640, 480)
        assert np.array_equal(got, want), (got, want)


def test_keypoint_geometry_semantics():
    cam = go.image_bounds(go.camera(bf=40.0, **TUM1), 640, 480)
    depth = np.zeros((480, 640), np.float32); depth[100, 200] = 2.0
    xy = np.float32([[200, 100], [0, 0], [639, 479]])
    un, kd, ur, cell = go.keypoint_geometry(xy, depth, cam)
    assert kd[0] == 2.0 and ur[0] == np.float32(un[0, 0] - np.float32(40.0) / np.float32(2.0)) and kd[1] == -1 and ur[2] == -1
    assert cell[0] >= 0 and np.all(cell < 64 * 48)
