// geom.cu -- the per-keypoint steps the reference's Frame constructor runs right after extraction, fused
// into ONE launch over the keypoints of a frame:
//   Frame::UndistortKeyPoints        src/Frame.cc:940-973   cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK)
//   Frame::ComputeStereoFromRGBD     src/Frame.cc:1177-1198 depth lookup at the (distorted) keypoint, virtual right coordinate
//   Frame::AssignFeaturesToGrid      src/Frame.cc:569-600   PosInGrid (:918-928) of the undistorted keypoint
// Third-party arithmetic: OpenCV's cvUndistortPointsInternal (calib3d/undistort.dispatch.cpp; un-vendored system package, README
// says 4.5.4): normalise, 5 fixed-point iterations of the radial / tangential model in double (the overload without a criteria
// argument uses TermCriteria(MAX_ITER, 5, 0.01), i.e. no epsilon test), re-project with P = mK, narrow to float.  Restated here
// operation by operation (no FMA contraction: __dmul_rn / __dadd_rn) and pinned against python-opencv 4.13 in
// tests/test_geom_oracle.py.  Byte work: 8 B in, 24 B out per keypoint + one depth texel.
#include <cuda_runtime.h>

#include "xfb_internal.h"

namespace xfb {

#ifdef __CUDA_ARCH__
#define XFB_DMUL(a, b) __dmul_rn((a), (b))
#define XFB_DADD(a, b) __dadd_rn((a), (b))
#define XFB_DSUB(a, b) __dsub_rn((a), (b))
#define XFB_DDIV(a, b) __ddiv_rn((a), (b))
#else
#define XFB_DMUL(a, b) ((a) * (b))
#define XFB_DADD(a, b) ((a) + (b))
#define XFB_DSUB(a, b) ((a) - (b))
#define XFB_DDIV(a, b) ((a) / (b))
#endif

// cvUndistortPointsInternal for one point, R = identity, P = K', distortion (k1, k2, p1, p2, k3), 5 iterations.
__host__ __device__ inline void undistort_point(float u_in, float v_in, const xfb_camera& cam, float* xo, float* yo) {
  const double fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy;
  const double k0 = cam.k1, k1 = cam.k2, k2 = cam.p1, k3 = cam.p2, k4 = cam.k3;
  const double ifx = XFB_DDIV(1.0, fx), ify = XFB_DDIV(1.0, fy);
  const double u = u_in, v = v_in;
  double x = XFB_DMUL(XFB_DSUB(u, cx), ifx), y = XFB_DMUL(XFB_DSUB(v, cy), ify);
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = XFB_DADD(XFB_DMUL(x, x), XFB_DMUL(y, y));
    // icdist = (1 + ((k7 r2 + k6) r2 + k5) r2) / (1 + ((k4 r2 + k1) r2 + k0) r2), k5 = k6 = k7 = 0
    const double den = XFB_DADD(1.0, XFB_DMUL(XFB_DADD(XFB_DMUL(XFB_DADD(XFB_DMUL(k4, r2), k1), r2), k0), r2));
    const double icdist = XFB_DDIV(1.0, den);
    if (icdist < 0) { x = XFB_DMUL(XFB_DSUB(u, cx), ifx); y = XFB_DMUL(XFB_DSUB(v, cy), ify); break; }
    // deltaX = 2 k2 x y + k3 (r2 + 2 x x)  (+ k8 r2 + k9 r2 r2 = 0);  deltaY = k2 (r2 + 2 y y) + 2 k3 x y
    const double deltaX = XFB_DADD(XFB_DMUL(XFB_DMUL(XFB_DMUL(2.0, k2), x), y), XFB_DMUL(k3, XFB_DADD(r2, XFB_DMUL(XFB_DMUL(2.0, x), x))));
    const double deltaY = XFB_DADD(XFB_DMUL(k2, XFB_DADD(r2, XFB_DMUL(XFB_DMUL(2.0, y), y))), XFB_DMUL(XFB_DMUL(XFB_DMUL(2.0, k3), x), y));
    x = XFB_DMUL(XFB_DSUB(x0, deltaX), icdist);
    y = XFB_DMUL(XFB_DSUB(y0, deltaY), icdist);
  }
  // xx = RR00 x + RR01 y + RR02 with RR = K' (R = I): fx x + 0 y + cx;  ww = 1 / (0 x + 0 y + 1) = 1
  *xo = (float)XFB_DADD(XFB_DMUL(fx, x), cx);
  *yo = (float)XFB_DADD(XFB_DMUL(fy, y), cy);
}

__global__ void __launch_bounds__(128) keypoint_geometry_kernel(const float* __restrict__ xy, int n, const float* __restrict__ depth, int h, int w,
                                                                int depth_stride, const xfb_camera cam, float* __restrict__ un_xy,
                                                                float* __restrict__ kp_depth, float* __restrict__ uright, int32_t* __restrict__ cell) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float u = xy[2 * i], v = xy[2 * i + 1];
  float xu = u, yu = v;
  if (cam.k1 != 0.0f) undistort_point(u, v, cam, &xu, &yu);   // mDistCoef.at<float>(0) == 0.0 -> mvKeysUn = mvKeys (src/Frame.cc:942-946)
  if (un_xy) { un_xy[2 * i] = xu; un_xy[2 * i + 1] = yu; }
  // ComputeStereoFromRGBD: d = imDepth.at<float>(v, u) (float -> int truncation), valid when d > 0
  float d_out = -1.0f, r_out = -1.0f;
  if (depth) {
    const int row = (int)v, col = (int)u;
    if (row >= 0 && row < h && col >= 0 && col < w) {
      const float d = depth[(size_t)row * depth_stride + col];
      if (d > 0.f) { d_out = d; r_out = __fsub_rn(xu, __fdiv_rn(cam.bf, d)); }
    }
  }
  if (kp_depth) kp_depth[i] = d_out;
  if (uright) uright[i] = r_out;
  // PosInGrid: round((x - mnMinX) * mfGridElementWidthInv), mfGridElementWidthInv = FRAME_GRID_COLS / (mnMaxX - mnMinX)
  if (cell) {
    const float wInv = __fdiv_rn(64.0f, __fsub_rn(cam.max_x, cam.min_x)), hInv = __fdiv_rn(48.0f, __fsub_rn(cam.max_y, cam.min_y));
    const int px = (int)roundf(__fmul_rn(__fsub_rn(xu, cam.min_x), wInv)), py = (int)roundf(__fmul_rn(__fsub_rn(yu, cam.min_y), hInv));
    cell[i] = (px < 0 || px >= 64 || py < 0 || py >= 48) ? -1 : px * 48 + py;   // mGrid[posX][posY]
  }
}

// Frame::ComputeImageBounds (src/Frame.cc:975-1003), host side: the four undistorted image corners.
void image_bounds_host(xfb_camera* cam, int w, int h) {
  if (cam->k1 == 0.0f) { cam->min_x = 0.f; cam->max_x = (float)w; cam->min_y = 0.f; cam->max_y = (float)h; return; }
  float x[4], y[4];
  const float cu[4] = {0.f, (float)w, 0.f, (float)w}, cv[4] = {0.f, 0.f, (float)h, (float)h};
  for (int i = 0; i < 4; ++i) undistort_point(cu[i], cv[i], *cam, &x[i], &y[i]);
  cam->min_x = x[0] < x[2] ? x[0] : x[2];
  cam->max_x = x[1] > x[3] ? x[1] : x[3];
  cam->min_y = y[0] < y[1] ? y[0] : y[1];
  cam->max_y = y[2] > y[3] ? y[2] : y[3];
}

cudaError_t launch_keypoint_geometry(Ctx* c, const float* d_xy, int n, const float* d_depth, int h, int w, int depth_stride, const xfb_camera& cam,
                                     float* d_un, float* d_depth_out, float* d_uright, int32_t* d_cell) {
  if (n <= 0) return cudaSuccess;
  keypoint_geometry_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(d_xy, n, d_depth, h, w, depth_stride, cam, d_un, d_depth_out, d_uright, d_cell);
  c->launches++;
  return cudaGetLastError();
}

}  // namespace xfb
