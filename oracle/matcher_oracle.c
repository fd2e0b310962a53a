/* matcher_oracle.c -- CPU restatement of the reference's descriptor matching primitives.
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product path (libxfeat_b200.so) never does.
 *
 * Plain C, single thread (the reference's matchers are single-threaded).  Each function cites the
 * reference file:line it follows.  src/ORBmatcher.cc cannot be compiled here (needs OpenCV, Eigen,
 * Sophus, DBoW2, Frame/KeyFrame/MapPoint -- SURVEY.md section 8c), so this is a "port" oracle.
 *
 * PARITY PINNING: the one third-party arithmetic in the path is OpenCV's
 * cv::norm(a, b, NORM_L2SQR) on CV_32F rows (un-vendored system libopencv, README says 4.5.4,
 * unpinned).  Its published algorithm for float inputs is normL2Sqr_<float,double>:
 * d = (float)a[i] - (float)b[i] in float, then accumulated as double s += (double)d * d; the
 * accumulation ORDER is build/SIMD dependent.  This oracle fixes sequential order i = 0..63.
 * tests/test_matcher_oracle.py pins it against cv2.norm (python-opencv 4.13 in this image) on
 * random unit-vector pairs; the int results agree except for ~1e-4 of pairs that sit within one
 * double-rounding of an integer boundary (the residual SURVEY.md 8c documents).  The reference has
 * no tests / golden vectors for the matcher.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define XF_DIM 64
#define GRID_COLS 64 /* FRAME_GRID_COLS, include/Frame.h:48 */
#define GRID_ROWS 48 /* FRAME_GRID_ROWS, include/Frame.h:47 */

/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:2242-2250 (USE_ORB unset branch):
 *   float normDist = cv::norm(a, b, cv::NORM_L2SQR); return static_cast<int>(normDist * 512); */
int mo_descriptor_distance(const float* a, const float* b) {
  double s = 0.0;
  for (int i = 0; i < XF_DIM; ++i) {
    float d = a[i] - b[i];
    s += (double)d * (double)d;
  }
  float nd = (float)s;
  return (int)(nd * 512.0f);
}

/* All-pairs DescriptorDistance: out[i*n2 + j] = dist(A_i, B_j). */
void mo_distance_matrix(const float* A, int n1, const float* B, int n2, int32_t* out) {
  for (int i = 0; i < n1; ++i)
    for (int j = 0; j < n2; ++j) out[(size_t)i * n2 + j] = mo_descriptor_distance(A + (size_t)i * XF_DIM, B + (size_t)j * XF_DIM);
}

/* Brute-force best / second-best scan with the update rule every live matcher uses
 * (e.g. src/ORBmatcher.cc:476-486, :884-894):
 *     if (dist < best) { second = best; best = dist; idx = j; } else if (dist < second) second = dist;
 * scanning j ascending, so the lowest index wins ties.  `init` is the initial best/second value
 * (256 in SearchByBoW :450-452, INT_MAX in SearchForInitialization :860-861).  groupA/groupB
 * (nullable) restrict candidates to equal group ids -- the vocabulary-node gating of
 * SearchByBoW/SearchForTriangulation (src/ORBmatcher.cc:430-436).  best_idx_rev[j] is the same
 * scan run column-wise (argmin over i, lowest i wins ties), for mutual-NN checks (the spec of the
 * commented-out ORBmatcher::match, src/ORBmatcher.cc:340-406, restated on the integer distance). */
void mo_bruteforce(const float* A, int n1, const float* B, int n2, const int32_t* groupA, const int32_t* groupB, int init,
                   int32_t* best_idx, int32_t* best_dist, int32_t* second_dist, int32_t* best_idx_rev, int32_t* best_dist_rev) {
  for (int j = 0; j < n2; ++j) { best_idx_rev[j] = -1; best_dist_rev[j] = init; }
  for (int i = 0; i < n1; ++i) {
    int b1 = init, b2 = init, bi = -1;
    for (int j = 0; j < n2; ++j) {
      if (groupA && groupB && groupA[i] != groupB[j]) continue;
      int d = mo_descriptor_distance(A + (size_t)i * XF_DIM, B + (size_t)j * XF_DIM);
      if (d < b1) { b2 = b1; b1 = d; bi = j; }
      else if (d < b2) { b2 = d; }
      if (d < best_dist_rev[j]) { best_dist_rev[j] = d; best_idx_rev[j] = i; }
    }
    best_idx[i] = bi; best_dist[i] = b1; second_dist[i] = b2;
  }
}

/* ---- Frame grid (src/Frame.cc:569-600 AssignFeaturesToGrid, :918-928 PosInGrid) -------------
 * Undistorted pinhole frame: mnMinX = 0, mnMaxX = cols, mnMinY = 0, mnMaxY = rows
 * (ComputeImageBounds with zero distortion), mfGridElementWidthInv = 64 / (mnMaxX - mnMinX). */
typedef struct {
  int* cell_start; /* [GRID_COLS*GRID_ROWS + 1] */
  int* items;      /* keypoint indices, cell-major, insertion (index) order inside a cell */
  float wInv, hInv;
} mo_grid;

static void grid_build(mo_grid* g, const float* kxy, int n, int img_w, int img_h) {
  const int nc = GRID_COLS * GRID_ROWS;
  g->wInv = (float)GRID_COLS / (float)(img_w - 0);
  g->hInv = (float)GRID_ROWS / (float)(img_h - 0);
  g->cell_start = (int*)calloc((size_t)nc + 1, sizeof(int));
  g->items = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int* cell_of = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) {
    int px = (int)roundf((kxy[2 * i] - 0.0f) * g->wInv);
    int py = (int)roundf((kxy[2 * i + 1] - 0.0f) * g->hInv);
    if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) { cell_of[i] = -1; continue; }
    cell_of[i] = px * GRID_ROWS + py; /* mGrid[ix][iy] */
    g->cell_start[cell_of[i] + 1]++;
  }
  for (int c = 0; c < nc; ++c) g->cell_start[c + 1] += g->cell_start[c];
  int* fill = (int*)calloc((size_t)nc, sizeof(int));
  for (int i = 0; i < n; ++i)
    if (cell_of[i] >= 0) g->items[g->cell_start[cell_of[i]] + fill[cell_of[i]]++] = i;
  free(fill);
  free(cell_of);
}
static void grid_free(mo_grid* g) { free(g->cell_start); free(g->items); }

/* Frame::GetFeaturesInArea, src/Frame.cc:850-916 (octave always 0 for XFeat => level checks pass).
 * Writes candidate indices in the reference's visiting order; returns the count. */
static int grid_area(const mo_grid* g, const float* kxy, float x, float y, float r, int* out) {
  int n = 0;
  int minCX = (int)floorf((x - 0.0f - r) * g->wInv); if (minCX < 0) minCX = 0;
  if (minCX >= GRID_COLS) return 0;
  int maxCX = (int)ceilf((x - 0.0f + r) * g->wInv); if (maxCX > GRID_COLS - 1) maxCX = GRID_COLS - 1;
  if (maxCX < 0) return 0;
  int minCY = (int)floorf((y - 0.0f - r) * g->hInv); if (minCY < 0) minCY = 0;
  if (minCY >= GRID_ROWS) return 0;
  int maxCY = (int)ceilf((y - 0.0f + r) * g->hInv); if (maxCY > GRID_ROWS - 1) maxCY = GRID_ROWS - 1;
  if (maxCY < 0) return 0;
  for (int ix = minCX; ix <= maxCX; ++ix)
    for (int iy = minCY; iy <= maxCY; ++iy) {
      int c = ix * GRID_ROWS + iy;
      for (int k = g->cell_start[c]; k < g->cell_start[c + 1]; ++k) {
        int j = g->items[k];
        float dx = kxy[2 * j] - x, dy = kxy[2 * j + 1] - y;
        if (fabsf(dx) < r && fabsf(dy) < r) out[n++] = j;
      }
    }
  return n;
}

/* ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:833-948, for XFeat keypoints
 * (octave 0, angle -1 => the rotation histogram puts every match in one bin and removes none).
 *   k1xy/k2xy : undistorted keypoint positions [n,2] of F1 / F2 (mvKeysUn)
 *   prev      : vbPrevMatched [n1,2], updated in place (:942-945)
 *   matches12 : out [n1], -1 = unmatched
 * th_low = ORBmatcher::TH_LOW (100 for XFeat, :35); ratio = mfNNratio (0.9 at Tracking.cc:2518). */
int mo_search_for_initialization(const float* D1, const float* k1xy, int n1, const float* D2, const float* k2xy, int n2,
                                 int img_w, int img_h, float* prev, int window, float ratio, int th_low, int32_t* matches12) {
  int nmatches = 0;
  mo_grid g;
  grid_build(&g, k2xy, n2, img_w, img_h);
  int* dist2 = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
  int* m21 = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
  int* cand = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
  for (int j = 0; j < n2; ++j) { dist2[j] = INT_MAX; m21[j] = -1; }
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  for (int i1 = 0; i1 < n1; ++i1) {
    int nc = grid_area(&g, k2xy, prev[2 * i1], prev[2 * i1 + 1], (float)window, cand);
    if (nc == 0) continue;
    int best = INT_MAX, best2 = INT_MAX, bidx = -1;
    for (int c = 0; c < nc; ++c) {
      int i2 = cand[c];
      int d = mo_descriptor_distance(D1 + (size_t)i1 * XF_DIM, D2 + (size_t)i2 * XF_DIM);
      if (dist2[i2] <= d) continue;
      if (d < best) { best2 = best; best = d; bidx = i2; }
      else if (d < best2) best2 = d;
    }
    if (best <= th_low) {
      if ((float)best < (float)best2 * ratio) {
        if (m21[bidx] >= 0) { matches12[m21[bidx]] = -1; nmatches--; }
        matches12[i1] = bidx; m21[bidx] = i1; dist2[bidx] = best; nmatches++;
      }
    }
  }
  for (int i1 = 0; i1 < n1; ++i1)
    if (matches12[i1] >= 0) { prev[2 * i1] = k2xy[2 * matches12[i1]]; prev[2 * i1 + 1] = k2xy[2 * matches12[i1] + 1]; }
  free(dist2); free(m21); free(cand);
  grid_free(&g);
  return nmatches;
}
