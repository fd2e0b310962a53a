"""CPU restatement of the per-keypoint frame geometry -- TEST INFRASTRUCTURE ONLY (tests/ and smoke()).

Follows Frame::UndistortKeyPoints (src/Frame.cc:940-973), Frame::ComputeImageBounds (:975-1003), Frame::ComputeStereoFromRGBD
(:1177-1198) and Frame::PosInGrid / AssignFeaturesToGrid (:918-928, :569-600).  The third-party arithmetic is OpenCV's
cv::undistortPoints (cvUndistortPointsInternal, calib3d/undistort.dispatch.cpp; un-vendored system libopencv-dev, README: 4.5.4):
normalise with 1/fx, five fixed-point iterations of the radial-tangential model in double (the overload without a criteria
argument uses TermCriteria(MAX_ITER, 5, 0.01): no epsilon test), re-project with P, narrow to float.
PINNED: tests/test_geom_oracle.py compares undistort() with cv2.undistortPoints of python-opencv 4.13 (bit-identical floats on
the TUM1 calibration and on random calibrations)."""
import numpy as np

GRID_COLS, GRID_ROWS = 64, 48          # FRAME_GRID_COLS / FRAME_GRID_ROWS, include/Frame.h:47-48


def camera(fx, fy, cx, cy, k1=0.0, k2=0.0, p1=0.0, p2=0.0, k3=0.0, bf=40.0):
    """float32[14]: fx fy cx cy k1 k2 p1 p2 k3 bf min_x min_y max_x max_y (the xfb_camera struct)."""
    return np.array([fx, fy, cx, cy, k1, k2, p1, p2, k3, bf, 0, 0, 0, 0], np.float32)


def undistort(xy, cam):
    """cv::undistortPoints(xy, K, dist, R = I, P = K) for float32 points [n,2] -> float32 [n,2]."""
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    if cam[4] == 0.0:                                   # mDistCoef.at<float>(0) == 0.0: mvKeysUn = mvKeys (:942-946)
        return xy.copy()
    fx, fy, cx, cy, k0, k1, k2, k3, k4 = [np.float64(v) for v in cam[:9]]      # k = (k1, k2, p1, p2, k3)
    ifx, ify = 1.0 / fx, 1.0 / fy
    u, v = xy[:, 0].astype(np.float64), xy[:, 1].astype(np.float64)
    x, y = (u - cx) * ifx, (v - cy) * ify
    x0, y0 = x.copy(), y.copy()
    done = np.zeros(len(x), bool)
    for _ in range(5):
        r2 = x * x + y * y
        icdist = 1.0 / (1.0 + ((k4 * r2 + k1) * r2 + k0) * r2)
        bad = (icdist < 0) & ~done
        deltaX = 2.0 * k2 * x * y + k3 * (r2 + 2.0 * x * x)
        deltaY = k2 * (r2 + 2.0 * y * y) + 2.0 * k3 * x * y
        xn, yn = (x0 - deltaX) * icdist, (y0 - deltaY) * icdist
        upd = ~done & ~bad
        x, y = np.where(upd, xn, x), np.where(upd, yn, y)
        x, y = np.where(bad, x0, x), np.where(bad, y0, y)   # icdist < 0: back to the normalised input, stop iterating
        done |= bad
    return np.stack([(fx * x + cx).astype(np.float32), (fy * y + cy).astype(np.float32)], 1)


def image_bounds(cam, w, h):
    """Frame::ComputeImageBounds -> cam with min_x, min_y, max_x, max_y filled."""
    cam = np.array(cam, np.float32, copy=True)
    if cam[4] == 0.0:
        cam[10:14] = [0.0, 0.0, w, h]
        return cam
    c = undistort(np.array([[0, 0], [w, 0], [0, h], [w, h]], np.float32), cam)
    cam[10] = min(c[0, 0], c[2, 0]); cam[12] = max(c[1, 0], c[3, 0])
    cam[11] = min(c[0, 1], c[1, 1]); cam[13] = max(c[2, 1], c[3, 1])
    return cam


def keypoint_geometry(xy, depth, cam):
    """(mvKeysUn.pt [n,2], mvDepth [n], mvuRight [n], grid cell posX * 48 + posY or -1 [n]) for keypoints xy [n,2]."""
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    un = undistort(xy, cam)
    n = len(xy)
    kd = np.full(n, -1.0, np.float32); ur = np.full(n, -1.0, np.float32)
    if depth is not None:
        h, w = depth.shape
        row, col = xy[:, 1].astype(np.int64), xy[:, 0].astype(np.int64)          # at<float>(v, u): float -> int truncation
        inside = (row >= 0) & (row < h) & (col >= 0) & (col < w)
        d = np.where(inside, depth[np.clip(row, 0, h - 1), np.clip(col, 0, w - 1)], np.float32(-1)).astype(np.float32)
        ok = d > 0
        kd[ok] = d[ok]
        ur[ok] = (un[ok, 0] - (np.float32(cam[9]) / d[ok]).astype(np.float32)).astype(np.float32)
    wInv = np.float32(GRID_COLS) / np.float32(cam[12] - cam[10]); hInv = np.float32(GRID_ROWS) / np.float32(cam[13] - cam[11])
    fx_, fy_ = ((un[:, 0] - cam[10]).astype(np.float32) * wInv).astype(np.float32), ((un[:, 1] - cam[11]).astype(np.float32) * hInv).astype(np.float32)
    rnd = lambda t: (np.sign(t) * np.floor(np.abs(t.astype(np.float64)) + 0.5)).astype(np.int64)   # C round(): half away from zero
    px, py = rnd(fx_), rnd(fy_)
    cell = np.where((px < 0) | (px >= GRID_COLS) | (py < 0) | (py >= GRID_ROWS), -1, px * GRID_ROWS + py).astype(np.int32)
    return un, kd, ur, cell
