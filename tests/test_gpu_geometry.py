"""xfb_keypoint_geometry (csrc/geom.cu: undistort + RGB-D depth / virtual right coordinate + grid cell in one launch) against
oracle/geom_oracle.py, which is pinned bit-for-bit against cv2.undistortPoints (tests/test_geom_oracle.py)."""
import numpy as np
import pytest

from oracle import geom_oracle as go
from xfeatslam_b200.frames import synthetic_frame

pytestmark = pytest.mark.gpu
TUM1 = dict(fx=517.306408, fy=516.469215, cx=318.643040, cy=255.313989, k1=0.262383, k2=-0.953104, p1=-0.005358, p2=0.002628, k3=1.163314)


@pytest.mark.parametrize("distorted", [True, False])
def test_keypoint_geometry_bit_exact(xfb_vga, distorted):
    from xfeatslam_b200 import capi
    cam = go.camera(bf=40.0, **TUM1) if distorted else go.camera(535.4, 539.2, 320.1, 247.6, bf=40.0)
    cam = capi.image_bounds(cam, 640, 480)
    assert np.array_equal(cam, go.image_bounds(cam, 640, 480))
    o = xfb_vga.extract(synthetic_frame(31, 480, 640), 2000)
    n = int(o["n_valid"])
    xy = np.concatenate([o["kpts"][:n], np.zeros((2000 - n, 2), np.float32)]).astype(np.float32)   # nfeatures rows, phantom (0,0) rows
    xy = np.concatenate([xy, np.float32([[0, 0], [639, 479], [639, 0], [0, 479], [320, 240]])])
    rng = np.random.RandomState(5)
    depth = (0.5 + 4.0 * rng.rand(480, 640)).astype(np.float32)
    depth[rng.rand(480, 640) < 0.2] = 0.0                                  # holes of the depth camera
    un, kd, ur, cell = xfb_vga.keypoint_geometry(xy, depth, cam)
    wun, wkd, wur, wcell = go.keypoint_geometry(xy, depth, cam)
    assert np.array_equal(un, wun) and np.array_equal(kd, wkd) and np.array_equal(ur, wur) and np.array_equal(cell, wcell)
    assert (kd > 0).sum() > 1000 and (cell >= 0).sum() > 1900
    if distorted:
        assert np.abs(un - xy).max() > 1.0                                 # the TUM1 lens moves corner points by pixels
    # monocular frame: no depth image
    un2, kd2, ur2, cell2 = xfb_vga.keypoint_geometry(xy, None, cam)
    assert np.array_equal(un2, wun) and np.all(kd2 == -1) and np.all(ur2 == -1) and np.array_equal(cell2, wcell)
