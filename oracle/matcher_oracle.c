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

/* ---- DBoW2::FeatureVector (thirdparty/DBoW2/DBoW2/FeatureVector.h: std::map<NodeId, std::vector<unsigned int>>) ----
 * node[i] = vocabulary node of feature i at `levelsup` (Frame::ComputeBoW, src/Frame.cc:931-938: levelsup = 4), or -1 when
 * the feature's word has zero weight and is therefore absent (TemplatedVocabulary::transform, TemplatedVocabulary.h:1183-1190:
 * `if(w > 0) { v.addWeight(id, w); fv.addFeature(nid, i_feature); }`).  Features of a node are stored in ascending feature
 * index (addFeature appends, i_feature ascends).  The matchers walk two maps in ascending NodeId and only act on equal keys
 * (src/ORBmatcher.cc:429-436, :579-586): a merge join. */
typedef struct { int n_nodes; int* node_id; int* start; int* feat; } mo_featvec;

static int cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return (x > y) - (x < y); }

static void featvec_build(mo_featvec* fv, const int32_t* node, int n) {
  int* ids = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  int m = 0;
  for (int i = 0; i < n; ++i) if (node[i] >= 0) ids[m++] = node[i];
  qsort(ids, (size_t)m, sizeof(int), cmp_int);
  int u = 0;
  for (int i = 0; i < m; ++i) if (i == 0 || ids[i] != ids[i - 1]) ids[u++] = ids[i];
  fv->n_nodes = u; fv->node_id = ids;
  fv->start = (int*)calloc((size_t)u + 1, sizeof(int));
  fv->feat = (int*)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
  for (int i = 0; i < n; ++i) {
    if (node[i] < 0) continue;
    int* p = (int*)bsearch(&node[i], ids, (size_t)u, sizeof(int), cmp_int);
    fv->start[(p - ids) + 1]++;
  }
  for (int k = 0; k < u; ++k) fv->start[k + 1] += fv->start[k];
  int* fill = (int*)calloc((size_t)(u > 0 ? u : 1), sizeof(int));
  for (int i = 0; i < n; ++i) {   /* ascending feature index inside a node */
    if (node[i] < 0) continue;
    int k = (int)((int*)bsearch(&node[i], ids, (size_t)u, sizeof(int), cmp_int) - ids);
    fv->feat[fv->start[k] + fill[k]++] = i;
  }
  free(fill);
}
static void featvec_free(mo_featvec* fv) { free(fv->node_id); free(fv->start); free(fv->feat); }

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), src/ORBmatcher.cc:408-610, F.Nleft == -1 branch.
 *   good_kf[i]   : vpMapPointsKF[i] != NULL && !isBad()   (:440-446)
 *   matches_f[j] : out, index of the KF feature whose MapPoint frame feature j received (vpMapPointMatches[j] = pMP), -1 = NULL
 * th_low = TH_LOW, ratio = mfNNratio (0.7 Tracking.cc:2754, 0.75 :3676).  XFeat keypoints all have angle -1, so every match
 * lands in rotation bin 0 and ComputeThreeMaxima removes nothing (:589-607). */
int mo_search_by_bow_kf_f(const float* Dkf, const int32_t* node_kf, const uint8_t* good_kf, int n_kf, const float* Df, const int32_t* node_f, int n_f,
                          float ratio, int th_low, int32_t* matches_f) {
  mo_featvec a, b;
  featvec_build(&a, node_kf, n_kf);
  featvec_build(&b, node_f, n_f);
  for (int j = 0; j < n_f; ++j) matches_f[j] = -1;
  int nmatches = 0, ia = 0, ib = 0;
  while (ia < a.n_nodes && ib < b.n_nodes) {
    if (a.node_id[ia] == b.node_id[ib]) {
      for (int p = a.start[ia]; p < a.start[ia + 1]; ++p) {
        const int realIdxKF = a.feat[p];
        if (!good_kf[realIdxKF]) continue;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int q = b.start[ib]; q < b.start[ib + 1]; ++q) {
          const int realIdxF = b.feat[q];
          if (matches_f[realIdxF] >= 0) continue;   /* :463 already claimed */
          const int dist = mo_descriptor_distance(Dkf + (size_t)realIdxKF * XF_DIM, Df + (size_t)realIdxF * XF_DIM);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= th_low) {
          if ((float)bestDist1 < ratio * (float)bestDist2) { matches_f[bestIdxF] = realIdxKF; nmatches++; }
        }
      }
      ia++; ib++;
    } else if (a.node_id[ia] < b.node_id[ib]) ia++;   /* lower_bound(Fit->first): first key >= ; stepping is equivalent in a merge */
    else ib++;
  }
  featvec_free(&a); featvec_free(&b);
  return nmatches;
}

/* ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&), src/ORBmatcher.cc:950-1090 (NLeft == -1).
 *   good1 / good2 : MapPoint present and not bad;  matches12[i1] : out, index of the KF2 feature (vpMatches12[idx1] = vpMapPoints2[bestIdx2])
 * Accept rule: bestDist1 < TH_LOW (strict, :1033) and ratio (0.9 at LoopClosing.cc:591). */
int mo_search_by_bow_kf_kf(const float* D1, const int32_t* node1, const uint8_t* good1, int n1, const float* D2, const int32_t* node2,
                           const uint8_t* good2, int n2, float ratio, int th_low, int32_t* matches12) {
  mo_featvec a, b;
  featvec_build(&a, node1, n1);
  featvec_build(&b, node2, n2);
  uint8_t* matched2 = (uint8_t*)calloc((size_t)(n2 > 0 ? n2 : 1), 1);
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  int nmatches = 0, ia = 0, ib = 0;
  while (ia < a.n_nodes && ib < b.n_nodes) {
    if (a.node_id[ia] == b.node_id[ib]) {
      for (int p = a.start[ia]; p < a.start[ia + 1]; ++p) {
        const int idx1 = a.feat[p];
        if (!good1[idx1]) continue;
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int q = b.start[ib]; q < b.start[ib + 1]; ++q) {
          const int idx2 = b.feat[q];
          if (matched2[idx2] || !good2[idx2]) continue;
          const int dist = mo_descriptor_distance(D1 + (size_t)idx1 * XF_DIM, D2 + (size_t)idx2 * XF_DIM);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 < th_low) {
          if ((float)bestDist1 < ratio * (float)bestDist2) { matches12[idx1] = bestIdx2; matched2[bestIdx2] = 1; nmatches++; }
        }
      }
      ia++; ib++;
    } else if (a.node_id[ia] < b.node_id[ib]) ia++;
    else ib++;
  }
  free(matched2);
  featvec_free(&a); featvec_free(&b);
  return nmatches;
}

/* ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:1092-1331, single pinhole camera (mpCamera2 == NULL, NLeft == -1).
 *   hasmp1 / hasmp2 : GetMapPoint(idx) != NULL (such features are skipped, :1159, :1189)
 *   stereo1 / stereo2 : mvuRight[idx] >= 0
 *   k1xy / k2xy : mvKeysUn positions;  ep : epipole of camera 1 in image 2 (:1105);  F12 : row-major fundamental matrix the
 *   reference builds inside Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:107-129: K1^-T [t12]x R12 K2^-1, Eigen
 *   float arithmetic -- supplied by the caller, Eigen is not in this image);  unc = mvLevelSigma2[kp2.octave] (octave 0: 1.0),
 *   scale0 = mvScaleFactors[0] = 1.
 * vbMatched2 is never set by the reference, so one KF2 feature may be taken by several KF1 features. */
int mo_search_for_triangulation(const float* D1, const int32_t* node1, const uint8_t* hasmp1, const uint8_t* stereo1, const float* k1xy, int n1,
                                const float* D2, const int32_t* node2, const uint8_t* hasmp2, const uint8_t* stereo2, const float* k2xy, int n2,
                                const float* F12, const float* ep, int only_stereo, int coarse, int th_low, float unc, float scale0,
                                int32_t* matches12) {
  mo_featvec a, b;
  featvec_build(&a, node1, n1);
  featvec_build(&b, node2, n2);
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  int nmatches = 0, ia = 0, ib = 0;
  while (ia < a.n_nodes && ib < b.n_nodes) {
    if (a.node_id[ia] == b.node_id[ib]) {
      for (int p = a.start[ia]; p < a.start[ia + 1]; ++p) {
        const int idx1 = a.feat[p];
        if (hasmp1[idx1]) continue;
        const int bStereo1 = stereo1[idx1];
        if (only_stereo && !bStereo1) continue;
        const float x1 = k1xy[2 * idx1], y1 = k1xy[2 * idx1 + 1];
        int bestDist = th_low, bestIdx2 = -1;
        for (int q = b.start[ib]; q < b.start[ib + 1]; ++q) {
          const int idx2 = b.feat[q];
          if (hasmp2[idx2]) continue;
          const int bStereo2 = stereo2[idx2];
          if (only_stereo && !bStereo2) continue;
          const int dist = mo_descriptor_distance(D1 + (size_t)idx1 * XF_DIM, D2 + (size_t)idx2 * XF_DIM);
          if (dist > th_low || dist > bestDist) continue;
          const float x2 = k2xy[2 * idx2], y2 = k2xy[2 * idx2 + 1];
          if (!bStereo1 && !bStereo2) {
            const float distex = ep[0] - x2, distey = ep[1] - y2;
            if (distex * distex + distey * distey < 100 * scale0) continue;
          }
          int ok = coarse;
          if (!ok) {   /* Pinhole::epipolarConstrain, src/CameraModels/Pinhole.cpp:114-128 */
            const float la = x1 * F12[0] + y1 * F12[3] + F12[6];
            const float lb = x1 * F12[1] + y1 * F12[4] + F12[7];
            const float lc = x1 * F12[2] + y1 * F12[5] + F12[8];
            const float num = la * x2 + lb * y2 + lc;
            const float den = la * la + lb * lb;
            if (den == 0) ok = 0;
            else { const float dsqr = num * num / den; ok = (double)dsqr < 3.84 * (double)unc; }
          }
          if (ok) { bestIdx2 = idx2; bestDist = dist; }
        }
        if (bestIdx2 >= 0) { matches12[idx1] = bestIdx2; nmatches++; }
      }
      ia++; ib++;
    } else if (a.node_id[ia] < b.node_id[ib]) ia++;
    else ib++;
  }
  featvec_free(&a); featvec_free(&b);
  return nmatches;
}

/* Frame::GetFeaturesInArea with level limits (src/Frame.cc:850-916) for keypoints that are all at octave 0. */
static int grid_area_levels(const mo_grid* g, const float* kxy, float x, float y, float r, int minLevel, int maxLevel, int* out) {
  const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  if (bCheckLevels && 0 < minLevel) return 0;               /* kpUn.octave (0) < minLevel */
  (void)maxLevel;                                            /* octave 0 > maxLevel is impossible for maxLevel >= 0 */
  return grid_area(g, kxy, x, y, r, out);
}

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints), src/ORBmatcher.cc:42-212,
 * monocular / RGB-D frames (F.Nleft == -1), the TrackLocalMap search.
 * Per map point m: in_view = mbTrackInView && !(bFarPoints && mTrackDepth > thFarPoints) && !isBad();  proj = (mTrackProjX,
 * mTrackProjY);  projxr = mTrackProjXR;  level = mnTrackScaleLevel;  viewcos = mTrackViewCos;  mp_obs = Observations() > 0.
 * Per frame feature: kxy = mvKeysUn, occupied = (F.mvpMapPoints[idx] && Observations() > 0), uright = mvuRight.
 * assign[idx] (out) = index of the map point written to F.mvpMapPoints[idx], -1 = untouched. */
int mo_search_by_projection(const float* Dmp, const uint8_t* in_view, const float* proj, const float* projxr, const int32_t* level,
                            const float* viewcos, const uint8_t* mp_obs, int n_mp, const float* Df, const float* kxy, const uint8_t* occupied_in,
                            const float* uright, int n_f, int img_w, int img_h, float th, float scale_factor, float ratio, int th_high,
                            int32_t* assign) {
  mo_grid g;
  grid_build(&g, kxy, n_f, img_w, img_h);
  int* cand = (int*)malloc(sizeof(int) * (size_t)(n_f > 0 ? n_f : 1));
  uint8_t* occupied = (uint8_t*)malloc((size_t)(n_f > 0 ? n_f : 1));
  for (int j = 0; j < n_f; ++j) { assign[j] = -1; occupied[j] = occupied_in[j]; }
  const int bFactor = th != 1.0f;
  int nmatches = 0;
  for (int m = 0; m < n_mp; ++m) {
    if (!in_view[m]) continue;
    const int lvl = level[m];
    float r = (viewcos[m] > 0.998f) ? 2.5f : 4.0f;          /* RadiusByViewingCos, :214-220 */
    if (bFactor) r *= th;
    float sf = 1.0f;                                         /* F.mvScaleFactors[lvl] = scaleFactor^lvl (XFextractor.cc:75-95) */
    for (int l = 0; l < lvl; ++l) sf *= scale_factor;
    const int nc = grid_area_levels(&g, kxy, proj[2 * m], proj[2 * m + 1], r * sf, lvl - 1, lvl, cand);
    if (nc == 0) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int c = 0; c < nc; ++c) {
      const int idx = cand[c];
      if (occupied[idx]) continue;
      if (uright[idx] > 0) {
        const float er = fabsf(projxr[m] - uright[idx]);
        if (er > r * sf) continue;
      }
      const int dist = mo_descriptor_distance(Dmp + (size_t)m * XF_DIM, Df + (size_t)idx * XF_DIM);
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = 0; bestIdx = idx; }
      else if (dist < bestDist2) { bestLevel2 = 0; bestDist2 = dist; }
    }
    if (bestDist <= th_high) {
      if (bestLevel == bestLevel2 && (float)bestDist > ratio * (float)bestDist2) continue;
      if (bestLevel != bestLevel2 || (float)bestDist <= ratio * (float)bestDist2) {
        assign[bestIdx] = m;
        occupied[bestIdx] = mp_obs[m];
        nmatches++;
      }
    }
  }
  free(cand); free(occupied);
  grid_free(&g);
  return nmatches;
}

/* MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:329-403, for `n_sets` map points at once: the observations of
 * set s are the rows offsets[s] .. offsets[s+1]-1 of D.  best[s] = the row (relative to the set) with the least median
 * distance to the others: Distances[i][i] = 0, median = sorted[0.5 * (N - 1)] (truncated), first minimum wins (:388-393). */
void mo_distinctive_descriptors(const float* D, const int32_t* offsets, int n_sets, int32_t* best) {
  for (int s = 0; s < n_sets; ++s) {
    const int o = offsets[s], N = offsets[s + 1] - offsets[s];
    best[s] = 0;
    if (N <= 0) { best[s] = -1; continue; }
    int* dist = (int*)malloc(sizeof(int) * (size_t)N * N);
    for (int i = 0; i < N; ++i) {
      dist[i * N + i] = 0;
      for (int j = i + 1; j < N; ++j) {
        const int d = mo_descriptor_distance(D + (size_t)(o + i) * XF_DIM, D + (size_t)(o + j) * XF_DIM);
        dist[i * N + j] = d; dist[j * N + i] = d;
      }
    }
    int BestMedian = INT_MAX, BestIdx = 0;
    int* row = (int*)malloc(sizeof(int) * (size_t)N);
    for (int i = 0; i < N; ++i) {
      memcpy(row, dist + (size_t)i * N, sizeof(int) * (size_t)N);
      qsort(row, (size_t)N, sizeof(int), cmp_int);
      const int median = row[(int)(0.5 * (N - 1))];
      if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    best[s] = BestIdx;
    free(row); free(dist);
  }
}

/* ---- DBoW2 vocabulary tree walk ----------------------------------------------------------------------------------
 * FORB::distance, thirdparty/DBoW2/DBoW2/FORB.cpp:81-101: Hamming distance over 8 int32 words (the bit-count is the parallel
 * bithack; the result equals popcount).  For XFeat frames `a` is a 1 x 64 CV_32F row, so the 8 words are the bit patterns of
 * its first 8 floats; `b` is the node's 32-byte ORB word. */
static int forb_distance(const uint32_t* pa, const uint32_t* pb) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t v = pa[i] ^ pb[i];
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

/* TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup), TemplatedVocabulary.h:1218-1260, for n features.
 * Tree as in xfb_vocab_load: node 0 = root, children of node i = child_index[child_start[i] .. child_start[i+1]) in
 * m_nodes[i].children order.  leaf[i] = final_id (the caller maps it to word_id / weight), nid[i] = node at level L - levelsup
 * (0 when L - levelsup <= 0; -1 where the reference would leave *nid unset because a leaf came first). */
void mo_bow_transform(const float* desc, int n, const uint8_t* node_desc, const int32_t* child_start, const int32_t* child_index, int L,
                      int levelsup, int32_t* leaf, int32_t* nid) {
  const int nid_level = L - levelsup;
  for (int i = 0; i < n; ++i) {
    uint32_t f[8];
    memcpy(f, desc + (size_t)i * XF_DIM, 32);
    int final_id = 0, current_level = 0;
    nid[i] = (nid_level <= 0) ? 0 : -1;
    do {
      ++current_level;
      const int c0 = child_start[final_id], c1 = child_start[final_id + 1];
      final_id = child_index[c0];
      uint32_t w[8];
      memcpy(w, node_desc + (size_t)final_id * 32, 32);
      double best_d = forb_distance(f, w);
      for (int c = c0 + 1; c < c1; ++c) {
        const int id = child_index[c];
        memcpy(w, node_desc + (size_t)id * 32, 32);
        const double d = forb_distance(f, w);
        if (d < best_d) { best_d = d; final_id = id; }
      }
      if (current_level == nid_level) nid[i] = final_id;
    } while (child_start[final_id + 1] > child_start[final_id]);
    leaf[i] = final_id;
  }
}

/* Frame::GetFeaturesInArea with explicit level limits for keypoints that are all at octave 0 (src/Frame.cc:850-916). */
static int grid_area_minmax(const mo_grid* g, const float* kxy, float x, float y, float r, int minLevel, int maxLevel, int* out) {
  const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  if (bCheckLevels) {
    if (0 < minLevel) return 0;                      /* kpUn.octave < minLevel */
    if (maxLevel >= 0 && 0 > maxLevel) return 0;     /* never true for octave 0 */
  }
  return grid_area(g, kxy, x, y, r, out);
}

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono), src/ORBmatcher.cc:1861-2072, the
 * TrackWithMotionModel search, CurrentFrame.Nleft == -1.  The caller projects the map points of LastFrame with the current pose
 * (:1888-1903, Sophus / Eigen arithmetic, not restated): per last-frame feature i
 *   valid[i]   : pMP && !mvbOutlier[i] && invzc >= 0 && the projection lies inside the image bounds
 *   uv[i]      : projected position; invzc[i] : 1 / depth in the current camera; octave[i] : LastFrame.mvKeys[i].octave (0 for XFeat)
 *   mp_obs[i]  : pMP->Observations() > 0
 * forward / backward : bForward / bBackward (:1875-1876).  assign[i2] (out) = last-frame index whose MapPoint is written to
 * CurrentFrame.mvpMapPoints[i2].  XFeat angles are all -1, so the rotation histogram removes nothing (:2047-2069). */
int mo_search_by_projection_frames(const float* Dlast, const uint8_t* valid, const float* uv, const float* invzc, const int32_t* octave,
                                   const uint8_t* mp_obs, int n_last, const float* Dcur, const float* kxy, const uint8_t* occupied_in,
                                   const float* uright, int n_cur, int img_w, int img_h, float th, float scale_factor, float mbf, int forward,
                                   int backward, int th_high, int32_t* assign) {
  mo_grid g;
  grid_build(&g, kxy, n_cur, img_w, img_h);
  int* cand = (int*)malloc(sizeof(int) * (size_t)(n_cur > 0 ? n_cur : 1));
  uint8_t* occupied = (uint8_t*)malloc((size_t)(n_cur > 0 ? n_cur : 1));
  for (int j = 0; j < n_cur; ++j) { assign[j] = -1; occupied[j] = occupied_in[j]; }
  int nmatches = 0;
  for (int i = 0; i < n_last; ++i) {
    if (!valid[i]) continue;
    const int nLastOctave = octave[i];
    float sf = 1.0f;
    for (int l = 0; l < nLastOctave; ++l) sf *= scale_factor;
    const float radius = th * sf;
    int nc;
    if (forward) nc = grid_area_minmax(&g, kxy, uv[2 * i], uv[2 * i + 1], radius, nLastOctave, -1, cand);
    else if (backward) nc = grid_area_minmax(&g, kxy, uv[2 * i], uv[2 * i + 1], radius, 0, nLastOctave, cand);
    else nc = grid_area_minmax(&g, kxy, uv[2 * i], uv[2 * i + 1], radius, nLastOctave - 1, nLastOctave + 1, cand);
    if (nc == 0) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int c = 0; c < nc; ++c) {
      const int i2 = cand[c];
      if (occupied[i2]) continue;
      if (uright[i2] > 0) {
        const float ur = uv[2 * i] - mbf * invzc[i];
        const float er = fabsf(ur - uright[i2]);
        if (er > radius) continue;
      }
      const int dist = mo_descriptor_distance(Dlast + (size_t)i * XF_DIM, Dcur + (size_t)i2 * XF_DIM);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    /* Reference: `if(bestDist<=TH_HIGH) CurrentFrame.mvpMapPoints[bestIdx2]=pMP;` (:1955-1958).  With XFeat TH_HIGH = 1000 > the
     * initial bestDist 256, so when every candidate was skipped or is >= 256 away the reference writes mvpMapPoints[-1] (undefined
     * behaviour, and counts a match).  Restated as "no match": bestIdx2 must be valid. */
    if (bestDist <= th_high && bestIdx2 >= 0) {
      assign[bestIdx2] = i;
      occupied[bestIdx2] = mp_obs[i];
      nmatches++;
    }
  }
  free(cand); free(occupied);
  grid_free(&g);
  return nmatches;
}

/* ORBmatcher::SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, vpPoints, vpMatched, th, ratioHamming), src/ORBmatcher.cc:612-717
 * (loop closing / merging; :719-831 is the same search with an extra vpMatchedKF bookkeeping array).  The geometric pre-checks
 * (:633-662: bad / already found, positive depth, inside the image, distance range, viewing angle) and the projection are the
 * caller's (Sophus / Eigen): valid[m], uv[m], radius[m] = th * mvScaleFactors[nPredictedLevel], level[m] = nPredictedLevel.
 * matched_in[idx] = (vpMatched[idx] != NULL) on entry.  assign[idx] (out) = map point written to vpMatched[idx], -1 = untouched.
 * All keypoints are at octave 0. */
int mo_search_by_projection_sim3(const float* Dmp, const uint8_t* valid, const float* uv, const float* radius, const int32_t* level, int n_mp,
                                 const float* Dkf, const float* kxy, const uint8_t* matched_in, int n_kf, int img_w, int img_h, int th_low,
                                 float ratio_hamming, int32_t* assign) {
  mo_grid g;
  grid_build(&g, kxy, n_kf, img_w, img_h);
  int* cand = (int*)malloc(sizeof(int) * (size_t)(n_kf > 0 ? n_kf : 1));
  uint8_t* matched = (uint8_t*)malloc((size_t)(n_kf > 0 ? n_kf : 1));
  for (int j = 0; j < n_kf; ++j) { assign[j] = -1; matched[j] = matched_in[j]; }
  int nmatches = 0;
  for (int m = 0; m < n_mp; ++m) {
    if (!valid[m]) continue;
    const int nPredictedLevel = level[m];
    const int nc = grid_area(&g, kxy, uv[2 * m], uv[2 * m + 1], radius[m], cand);   /* GetFeaturesInArea(u, v, radius): no level limits */
    if (nc == 0) continue;
    int bestDist = 256, bestIdx = -1;
    for (int c = 0; c < nc; ++c) {
      const int idx = cand[c];
      if (matched[idx]) continue;
      const int kpLevel = 0;
      if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
      const int dist = mo_descriptor_distance(Dmp + (size_t)m * XF_DIM, Dkf + (size_t)idx * XF_DIM);
      if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
    if ((float)bestDist <= (float)th_low * ratio_hamming) { assign[bestIdx] = m; matched[bestIdx] = 1; nmatches++; }
  }
  free(cand); free(matched);
  grid_free(&g);
  return nmatches;
}

/* The candidate search of ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th, bRight = false), src/ORBmatcher.cc:1333-1523 (:1413-1479): per
 * map point the closest keypoint in the window that passes the level filter and the chi-square reprojection test.  The pre-checks
 * (:1363-1408) and the Replace / AddObservation decisions on the result (:1481-1503, `if(bestDist<=TH_LOW)`) stay with the caller.
 *   valid, uv, ur (= u - bf * invz), radius, level : per map point;  uright = pKF->mvuRight;  inv_sigma2_0 = mvInvLevelSigma2[0]
 * best_idx / best_dist (out): -1 / 256 where nothing qualified. */
void mo_fuse_search(const float* Dmp, const uint8_t* valid, const float* uv, const float* ur, const float* radius, const int32_t* level, int n_mp,
                    const float* Dkf, const float* kxy, const float* uright, int n_kf, int img_w, int img_h, float inv_sigma2_0, int32_t* best_idx,
                    int32_t* best_dist) {
  mo_grid g;
  grid_build(&g, kxy, n_kf, img_w, img_h);
  int* cand = (int*)malloc(sizeof(int) * (size_t)(n_kf > 0 ? n_kf : 1));
  for (int m = 0; m < n_mp; ++m) {
    best_idx[m] = -1; best_dist[m] = 256;
    if (!valid[m]) continue;
    const int nPredictedLevel = level[m];
    const int nc = grid_area(&g, kxy, uv[2 * m], uv[2 * m + 1], radius[m], cand);
    int bestDist = 256, bestIdx = -1;
    for (int c = 0; c < nc; ++c) {
      const int idx = cand[c];
      const int kpLevel = 0;
      if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
      const float kpx = kxy[2 * idx], kpy = kxy[2 * idx + 1];
      if (uright[idx] >= 0) {
        const float ex = uv[2 * m] - kpx, ey = uv[2 * m + 1] - kpy, er = ur[m] - uright[idx];
        const float e2 = ex * ex + ey * ey + er * er;
        if ((double)(e2 * inv_sigma2_0) > 7.8) continue;
      } else {
        const float ex = uv[2 * m] - kpx, ey = uv[2 * m + 1] - kpy;
        const float e2 = ex * ex + ey * ey;
        if ((double)(e2 * inv_sigma2_0) > 5.99) continue;
      }
      const int dist = mo_descriptor_distance(Dmp + (size_t)m * XF_DIM, Dkf + (size_t)idx * XF_DIM);
      if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
    best_idx[m] = bestIdx; best_dist[m] = bestDist;
  }
  free(cand);
  grid_free(&g);
}

/* One direction of ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1690-1760 / :1762-1833): closest keypoint of the other keyframe in
 * the window, level filter, bestDist starts at INT_MAX, accepted when <= TH_HIGH. */
static void sim3_one_way(const float* Dmp, const uint8_t* valid, const float* uv, const float* radius, const int32_t* level, int n_mp,
                         const float* Dkf, const float* kxy, int n_kf, int img_w, int img_h, int th_high, int* match) {
  mo_grid g;
  grid_build(&g, kxy, n_kf, img_w, img_h);
  int* cand = (int*)malloc(sizeof(int) * (size_t)(n_kf > 0 ? n_kf : 1));
  for (int i = 0; i < n_mp; ++i) {
    match[i] = -1;
    if (!valid[i]) continue;
    const int nc = grid_area(&g, kxy, uv[2 * i], uv[2 * i + 1], radius[i], cand);
    if (nc == 0) continue;
    int bestDist = INT_MAX, bestIdx = -1;
    for (int c = 0; c < nc; ++c) {
      const int idx = cand[c];
      if (0 < level[i] - 1 || 0 > level[i]) continue;   /* kp.octave (0) < nPredictedLevel - 1 || > nPredictedLevel */
      const int dist = mo_descriptor_distance(Dmp + (size_t)i * XF_DIM, Dkf + (size_t)idx * XF_DIM);
      if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
    if (bestDist <= th_high) match[i] = bestIdx;
  }
  free(cand);
  grid_free(&g);
}

/* ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th), src/ORBmatcher.cc:1642-1859.  Per feature of KF1 (resp. KF2): valid =
 * has a good MapPoint, not already matched, positive depth, inside the other image, distance in range (:1692-1722 / :1764-1794);
 * uv / radius / level = its projection into the other keyframe with S21 (resp. S12); Dmp1 / Dmp2 rows = pMP->GetDescriptor().
 * matches12[i1] (out) = idx2 where both directions agree (:1835-1850), else -1; returns nFound. */
int mo_search_by_sim3(const float* Dmp1, const uint8_t* valid1, const float* uv1, const float* radius1, const int32_t* level1, int n1,
                      const float* D2, const float* k2xy, const float* Dmp2, const uint8_t* valid2, const float* uv2, const float* radius2,
                      const int32_t* level2, int n2, const float* D1, const float* k1xy, int img_w, int img_h, int th_high, int32_t* matches12) {
  int* m1 = (int*)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1));
  int* m2 = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
  sim3_one_way(Dmp1, valid1, uv1, radius1, level1, n1, D2, k2xy, n2, img_w, img_h, th_high, m1);
  sim3_one_way(Dmp2, valid2, uv2, radius2, level2, n2, D1, k1xy, n1, img_w, img_h, th_high, m2);
  int nFound = 0;
  for (int i1 = 0; i1 < n1; ++i1) {
    matches12[i1] = -1;
    const int idx2 = m1[i1];
    if (idx2 >= 0 && m2[idx2] == i1) { matches12[i1] = idx2; nFound++; }
  }
  free(m1); free(m2);
  return nFound;
}

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:2074-2190
 * (relocalisation).  Per map point of the keyframe: valid = pMP && !isBad() && !sAlreadyFound.count(pMP) && projection inside the
 * image bounds && distance in range (:2095-2121); uv, radius = th * mvScaleFactors[nPredictedLevel], level = nPredictedLevel.
 * occupied_in[i2] = (CurrentFrame.mvpMapPoints[i2] != NULL).  assign[i2] (out) = map point written there.  All octaves are 0 and all
 * angles -1 (the rotation histogram removes nothing, :2165-2187). */
int mo_search_by_projection_reloc(const float* Dmp, const uint8_t* valid, const float* uv, const float* radius, const int32_t* level, int n_mp,
                                  const float* Dcur, const float* kxy, const uint8_t* occupied_in, int n_cur, int img_w, int img_h, int orb_dist,
                                  int32_t* assign) {
  mo_grid g;
  grid_build(&g, kxy, n_cur, img_w, img_h);
  int* cand = (int*)malloc(sizeof(int) * (size_t)(n_cur > 0 ? n_cur : 1));
  uint8_t* occupied = (uint8_t*)malloc((size_t)(n_cur > 0 ? n_cur : 1));
  for (int j = 0; j < n_cur; ++j) { assign[j] = -1; occupied[j] = occupied_in[j]; }
  int nmatches = 0;
  for (int m = 0; m < n_mp; ++m) {
    if (!valid[m]) continue;
    const int nPredictedLevel = level[m];
    const int nc = grid_area_minmax(&g, kxy, uv[2 * m], uv[2 * m + 1], radius[m], nPredictedLevel - 1, nPredictedLevel + 1, cand);
    if (nc == 0) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (int c = 0; c < nc; ++c) {
      const int i2 = cand[c];
      if (occupied[i2]) continue;
      const int dist = mo_descriptor_distance(Dmp + (size_t)m * XF_DIM, Dcur + (size_t)i2 * XF_DIM);
      if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
    }
    if (bestDist <= orb_dist && bestIdx2 >= 0) {   /* (bestIdx2 >= 0 matters only for ORBdist >= 256: see mo_search_by_projection_frames) */
      assign[bestIdx2] = m; occupied[bestIdx2] = 1; nmatches++;
    }
  }
  free(cand); free(occupied);
  grid_free(&g);
  return nmatches;
}
