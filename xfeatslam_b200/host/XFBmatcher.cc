// XFBmatcher.cc -- see XFBmatcher.h.  Host control flow in C++, distances from libxfeat_b200.so.
#include "XFBmatcher.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace ORB_SLAM3 {

namespace {
const int kGridCols = 64, kGridRows = 48;   // FRAME_GRID_COLS / FRAME_GRID_ROWS, include/Frame.h:47-48

std::vector<float> packed_rows(const cv::Mat& m) {
  if (m.type() != CV_32F || m.cols != XFB_DESC_DIM) throw std::invalid_argument("XFBmatcher: descriptors must be CV_32F with 64 columns");
  std::vector<float> v(static_cast<size_t>(m.rows) * XFB_DESC_DIM);
  for (int r = 0; r < m.rows; ++r) std::memcpy(v.data() + static_cast<size_t>(r) * XFB_DESC_DIM, m.ptr<float>(r), XFB_DESC_DIM * sizeof(float));
  return v;
}

// Frame::AssignFeaturesToGrid + PosInGrid (src/Frame.cc:569-600, :918-928)
struct Grid {
  std::vector<std::vector<size_t>> cell;   // [ix * rows + iy]
  float wInv, hInv, minX, minY;
  Grid(const std::vector<cv::KeyPoint>& keys, float mnX, float mnY, float mxX, float mxY)
      : cell(static_cast<size_t>(kGridCols) * kGridRows), wInv(static_cast<float>(kGridCols) / (mxX - mnX)),
        hInv(static_cast<float>(kGridRows) / (mxY - mnY)), minX(mnX), minY(mnY) {
    for (size_t i = 0; i < keys.size(); ++i) {
      const int px = static_cast<int>(std::round((keys[i].pt.x - minX) * wInv));
      const int py = static_cast<int>(std::round((keys[i].pt.y - minY) * hInv));
      if (px < 0 || px >= kGridCols || py < 0 || py >= kGridRows) continue;
      cell[static_cast<size_t>(px) * kGridRows + py].push_back(i);
    }
  }
  // Frame::GetFeaturesInArea (src/Frame.cc:850-916), minLevel = maxLevel = 0 (XFeat keypoints are octave 0)
  std::vector<size_t> area(const std::vector<cv::KeyPoint>& keys, float x, float y, float r, int minLevel = -1, int maxLevel = -1) const {
    std::vector<size_t> out;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    const int nMinCellX = std::max(0, static_cast<int>(std::floor((x - minX - r) * wInv)));
    if (nMinCellX >= kGridCols) return out;
    const int nMaxCellX = std::min(kGridCols - 1, static_cast<int>(std::ceil((x - minX + r) * wInv)));
    if (nMaxCellX < 0) return out;
    const int nMinCellY = std::max(0, static_cast<int>(std::floor((y - minY - r) * hInv)));
    if (nMinCellY >= kGridRows) return out;
    const int nMaxCellY = std::min(kGridRows - 1, static_cast<int>(std::ceil((y - minY + r) * hInv)));
    if (nMaxCellY < 0) return out;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
      for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
        for (size_t j : cell[static_cast<size_t>(ix) * kGridRows + iy]) {
          if (bCheckLevels) {
            if (keys[j].octave < minLevel) continue;
            if (maxLevel >= 0 && keys[j].octave > maxLevel) continue;
          }
          if (std::fabs(keys[j].pt.x - x) < r && std::fabs(keys[j].pt.y - y) < r) out.push_back(j);
        }
    return out;
  }
};
}  // namespace

XFBmatcher::XFBmatcher(xfb_ctx* ctx, float nnratio, bool checkOri) : ctx_(ctx), mfNNratio(nnratio), mbCheckOrientation(checkOri) {
  if (!ctx) throw std::invalid_argument("XFBmatcher: null xfb_ctx (no CPU fallback)");
}

XFBmatcher::DistanceTable XFBmatcher::ComputeDistances(const cv::Mat& desc1, const cv::Mat& desc2) const {
  DistanceTable t;
  t.n1_ = desc1.rows; t.n2_ = desc2.rows;
  t.d_.resize(static_cast<size_t>(t.n1_) * t.n2_);
  if (t.n1_ == 0 || t.n2_ == 0) return t;
  const std::vector<float> a = packed_rows(desc1), b = packed_rows(desc2);
  if (xfb_distance_matrix(ctx_, a.data(), t.n1_, b.data(), t.n2_, t.d_.data()) != XFB_OK)
    throw std::runtime_error(std::string("xfb_distance_matrix: ") + xfb_last_error(ctx_));
  return t;
}

void XFBmatcher::match(cv::Mat d1, cv::Mat d2, std::vector<cv::DMatch>& matches) const {
  matches.clear();
  if (d1.rows == 0 || d2.rows == 0) return;
  const std::vector<float> a = packed_rows(d1), b = packed_rows(d2);
  std::vector<int32_t> bi(d1.rows), bd(d1.rows), ri(d2.rows);
  if (xfb_match(ctx_, a.data(), d1.rows, b.data(), d2.rows, nullptr, nullptr, INT_MAX, bi.data(), bd.data(), nullptr, ri.data(), nullptr) != XFB_OK)
    throw std::runtime_error(std::string("xfb_match: ") + xfb_last_error(ctx_));
  for (int i = 0; i < d1.rows; ++i)
    if (bi[i] >= 0 && ri[bi[i]] == i)                                           // mutual: match21[match12[i]] == i
      matches.emplace_back(i, bi[i], std::sqrt(static_cast<float>(bd[i]) / 512.0f));   // sqrt(2 (1 - cos)) == ||a - b||
}

int XFBmatcher::SearchForInitialization(const std::vector<cv::KeyPoint>& keys1, const cv::Mat& desc1, const std::vector<cv::KeyPoint>& keys2,
                                        const cv::Mat& desc2, float minX, float minY, float maxX, float maxY,
                                        std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize) const {
  int nmatches = 0;
  vnMatches12 = std::vector<int>(keys1.size(), -1);
  const DistanceTable D = ComputeDistances(desc1, desc2);
  const Grid grid(keys2, minX, minY, maxX, maxY);
  std::vector<int> vMatchedDistance(keys2.size(), INT_MAX);
  std::vector<int> vnMatches21(keys2.size(), -1);
  for (size_t i1 = 0, iend1 = keys1.size(); i1 < iend1; i1++) {
    if (keys1[i1].octave > 0) continue;
    const std::vector<size_t> vIndices2 = grid.area(keys2, vbPrevMatched[i1].x, vbPrevMatched[i1].y, static_cast<float>(windowSize), keys1[i1].octave, keys1[i1].octave);
    if (vIndices2.empty()) continue;
    int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
    for (size_t i2 : vIndices2) {
      const int dist = D(static_cast<int>(i1), static_cast<int>(i2));
      if (vMatchedDistance[i2] <= dist) continue;
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = static_cast<int>(i2); }
      else if (dist < bestDist2) bestDist2 = dist;
    }
    if (bestDist <= TH_LOW && bestDist < static_cast<float>(bestDist2) * mfNNratio) {
      if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
      vnMatches12[i1] = bestIdx2;
      vnMatches21[bestIdx2] = static_cast<int>(i1);
      vMatchedDistance[bestIdx2] = bestDist;
      nmatches++;
      // rotation histogram (:904-914): every XFeat keypoint has angle -1 => one bin, nothing is ever removed
    }
  }
  for (size_t i1 = 0; i1 < vnMatches12.size(); i1++)
    if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = keys2[vnMatches12[i1]].pt;
  return nmatches;
}

std::vector<int32_t> XFBmatcher::PairDistances(const cv::Mat& desc1, const cv::Mat& desc2, const std::vector<int32_t>& i1,
                                               const std::vector<int32_t>& i2) const {
  std::vector<int32_t> d(i1.size());
  if (i1.empty()) return d;
  const std::vector<float> a = packed_rows(desc1), b = packed_rows(desc2);
  if (xfb_distance_pairs(ctx_, a.data(), desc1.rows, b.data(), desc2.rows, i1.data(), i2.data(), static_cast<int>(i1.size()), d.data()) != XFB_OK)
    throw std::runtime_error(std::string("xfb_distance_pairs: ") + xfb_last_error(ctx_));
  return d;
}

namespace {
// The merge join every node-gated matcher runs over two DBoW2::FeatureVector maps (src/ORBmatcher.cc:429-436, :579-586):
// calls f(indices1, indices2) for every node id present in both, in ascending node id.
template <class Fn>
void for_shared_nodes(const XFBmatcher::FeatureVector& v1, const XFBmatcher::FeatureVector& v2, Fn f) {
  XFBmatcher::FeatureVector::const_iterator it1 = v1.begin(), it2 = v2.begin(), end1 = v1.end(), end2 = v2.end();
  while (it1 != end1 && it2 != end2) {
    if (it1->first == it2->first) { f(it1->second, it2->second); ++it1; ++it2; }
    else if (it1->first < it2->first) it1 = v1.lower_bound(it2->first);
    else it2 = v2.lower_bound(it1->first);
  }
}
}  // namespace

int XFBmatcher::SearchByBoW(const cv::Mat& descKF, const FeatureVector& vFeatVecKF, const std::vector<bool>& vbGoodMapPointKF, const cv::Mat& descF,
                            const FeatureVector& vFeatVecF, std::vector<int>& vnMatchesF) const {
  // 1. every (KF feature with a good MapPoint, F feature of the same node) pair, in the reference's visiting order
  std::vector<int32_t> p1, p2;
  for_shared_nodes(vFeatVecKF, vFeatVecF, [&](const std::vector<unsigned int>& vIndicesKF, const std::vector<unsigned int>& vIndicesF) {
    for (unsigned int realIdxKF : vIndicesKF) {
      if (!vbGoodMapPointKF[realIdxKF]) continue;
      for (unsigned int realIdxF : vIndicesF) { p1.push_back(static_cast<int32_t>(realIdxKF)); p2.push_back(static_cast<int32_t>(realIdxF)); }
    }
  });
  const std::vector<int32_t> dist = PairDistances(descKF, descF, p1, p2);
  // 2. the reference's loop (:436-574) over those distances
  vnMatchesF = std::vector<int>(descF.rows, -1);
  int nmatches = 0;
  size_t cur = 0;
  for_shared_nodes(vFeatVecKF, vFeatVecF, [&](const std::vector<unsigned int>& vIndicesKF, const std::vector<unsigned int>& vIndicesF) {
    for (unsigned int realIdxKF : vIndicesKF) {
      if (!vbGoodMapPointKF[realIdxKF]) continue;
      int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
      for (unsigned int realIdxF : vIndicesF) {
        const int d = dist[cur++];
        if (vnMatchesF[realIdxF] >= 0) continue;   // :463 vpMapPointMatches[realIdxF] already set
        if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdxF = static_cast<int>(realIdxF); }
        else if (d < bestDist2) bestDist2 = d;
      }
      if (bestDist1 <= TH_LOW && static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {
        vnMatchesF[bestIdxF] = static_cast<int>(realIdxKF);
        nmatches++;
        // rotation histogram (:523-538, :589-607): every XFeat keypoint has angle -1 => bin 0 only, nothing is removed
      }
    }
  });
  return nmatches;
}

int XFBmatcher::SearchByBoW(const cv::Mat& desc1, const FeatureVector& vFeatVec1, const std::vector<bool>& vbGoodMapPoint1, const cv::Mat& desc2,
                            const FeatureVector& vFeatVec2, const std::vector<bool>& vbGoodMapPoint2, std::vector<int>& vnMatches12) const {
  std::vector<int32_t> p1, p2;
  for_shared_nodes(vFeatVec1, vFeatVec2, [&](const std::vector<unsigned int>& f1, const std::vector<unsigned int>& f2) {
    for (unsigned int idx1 : f1) {
      if (!vbGoodMapPoint1[idx1]) continue;
      for (unsigned int idx2 : f2) {
        if (!vbGoodMapPoint2[idx2]) continue;
        p1.push_back(static_cast<int32_t>(idx1)); p2.push_back(static_cast<int32_t>(idx2));
      }
    }
  });
  const std::vector<int32_t> dist = PairDistances(desc1, desc2, p1, p2);
  vnMatches12 = std::vector<int>(desc1.rows, -1);
  std::vector<bool> vbMatched2(desc2.rows, false);
  int nmatches = 0;
  size_t cur = 0;
  for_shared_nodes(vFeatVec1, vFeatVec2, [&](const std::vector<unsigned int>& f1, const std::vector<unsigned int>& f2) {
    for (unsigned int idx1 : f1) {
      if (!vbGoodMapPoint1[idx1]) continue;
      int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
      for (unsigned int idx2 : f2) {
        if (!vbGoodMapPoint2[idx2]) continue;
        const int d = dist[cur++];
        if (vbMatched2[idx2]) continue;
        if (d < bestDist1) { bestDist2 = bestDist1; bestDist1 = d; bestIdx2 = static_cast<int>(idx2); }
        else if (d < bestDist2) bestDist2 = d;
      }
      if (bestDist1 < TH_LOW && static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {   // strict '<' at :1033
        vnMatches12[idx1] = bestIdx2;
        vbMatched2[bestIdx2] = true;
        nmatches++;
      }
    }
  });
  return nmatches;
}

int XFBmatcher::SearchForTriangulation(const cv::Mat& desc1, const FeatureVector& vFeatVec1, const std::vector<bool>& vbHasMapPoint1,
                                       const std::vector<bool>& vbStereo1, const std::vector<cv::KeyPoint>& vKeysUn1, const cv::Mat& desc2,
                                       const FeatureVector& vFeatVec2, const std::vector<bool>& vbHasMapPoint2, const std::vector<bool>& vbStereo2,
                                       const std::vector<cv::KeyPoint>& vKeysUn2, const float F12[9], const cv::Point2f& ep,
                                       std::vector<std::pair<size_t, size_t> >& vMatchedPairs, bool bOnlyStereo, bool bCoarse, float sigma2Level0,
                                       float scaleFactor0) const {
  auto skip1 = [&](unsigned int idx1) { return vbHasMapPoint1[idx1] || (bOnlyStereo && !vbStereo1[idx1]); };
  auto skip2 = [&](unsigned int idx2) { return vbHasMapPoint2[idx2] || (bOnlyStereo && !vbStereo2[idx2]); };
  std::vector<int32_t> p1, p2;
  for_shared_nodes(vFeatVec1, vFeatVec2, [&](const std::vector<unsigned int>& f1, const std::vector<unsigned int>& f2) {
    for (unsigned int idx1 : f1) {
      if (skip1(idx1)) continue;
      for (unsigned int idx2 : f2) {
        if (skip2(idx2)) continue;
        p1.push_back(static_cast<int32_t>(idx1)); p2.push_back(static_cast<int32_t>(idx2));
      }
    }
  });
  const std::vector<int32_t> dist = PairDistances(desc1, desc2, p1, p2);
  std::vector<int> vMatches12(desc1.rows, -1);
  int nmatches = 0;
  size_t cur = 0;
  for_shared_nodes(vFeatVec1, vFeatVec2, [&](const std::vector<unsigned int>& f1, const std::vector<unsigned int>& f2) {
    for (unsigned int idx1 : f1) {
      if (skip1(idx1)) continue;
      const bool bStereo1 = vbStereo1[idx1];
      const cv::KeyPoint& kp1 = vKeysUn1[idx1];
      int bestDist = TH_LOW, bestIdx2 = -1;
      for (unsigned int idx2 : f2) {
        if (skip2(idx2)) continue;
        const int d = dist[cur++];
        if (d > TH_LOW || d > bestDist) continue;
        const cv::KeyPoint& kp2 = vKeysUn2[idx2];
        if (!bStereo1 && !vbStereo2[idx2]) {
          const float distex = ep.x - kp2.pt.x, distey = ep.y - kp2.pt.y;
          if (distex * distex + distey * distey < 100 * scaleFactor0) continue;
        }
        bool ok = bCoarse;
        if (!ok) {   // Pinhole::epipolarConstrain, src/CameraModels/Pinhole.cpp:114-128
          const float a = kp1.pt.x * F12[0] + kp1.pt.y * F12[3] + F12[6];
          const float b = kp1.pt.x * F12[1] + kp1.pt.y * F12[4] + F12[7];
          const float c = kp1.pt.x * F12[2] + kp1.pt.y * F12[5] + F12[8];
          const float num = a * kp2.pt.x + b * kp2.pt.y + c;
          const float den = a * a + b * b;
          if (den == 0) ok = false;
          else { const float dsqr = num * num / den; ok = dsqr < 3.84 * sigma2Level0; }
        }
        if (ok) { bestIdx2 = static_cast<int>(idx2); bestDist = d; }
      }
      if (bestIdx2 >= 0) { vMatches12[idx1] = bestIdx2; nmatches++; }
    }
  });
  vMatchedPairs.clear();
  vMatchedPairs.reserve(nmatches);
  for (size_t i = 0, iend = vMatches12.size(); i < iend; i++) {
    if (vMatches12[i] < 0) continue;
    vMatchedPairs.push_back(std::make_pair(i, static_cast<size_t>(vMatches12[i])));
  }
  return nmatches;
}

int XFBmatcher::SearchByProjection(const std::vector<ProjectedPoint>& vPoints, const cv::Mat& descMP, const std::vector<cv::KeyPoint>& vKeysUnF,
                                   const cv::Mat& descF, const std::vector<bool>& vbOccupiedF, const std::vector<float>& vuRightF, float minX,
                                   float minY, float maxX, float maxY, float scaleFactor, float th, std::vector<int>& vnAssignedF) const {
  const Grid grid(vKeysUnF, minX, minY, maxX, maxY);
  const bool bFactor = th != 1.0;
  // 1. candidate windows (host, like the reference) -> one pair list
  std::vector<std::vector<size_t> > cand(vPoints.size());
  std::vector<float> radius(vPoints.size(), 0.f);
  std::vector<int32_t> p1, p2;
  for (size_t iMP = 0; iMP < vPoints.size(); iMP++) {
    const ProjectedPoint& mp = vPoints[iMP];
    if (!mp.inView) continue;
    float r = (mp.viewCos > 0.998) ? 2.5f : 4.0f;   // RadiusByViewingCos, :214-220
    if (bFactor) r *= th;
    float sf = 1.0f;                                 // F.mvScaleFactors[nPredictedLevel]
    for (int l = 0; l < mp.scaleLevel; ++l) sf *= scaleFactor;
    radius[iMP] = r * sf;
    cand[iMP] = grid.area(vKeysUnF, mp.projX, mp.projY, r * sf, mp.scaleLevel - 1, mp.scaleLevel);
    for (size_t idx : cand[iMP]) { p1.push_back(static_cast<int32_t>(iMP)); p2.push_back(static_cast<int32_t>(idx)); }
  }
  const std::vector<int32_t> dist = PairDistances(descMP, descF, p1, p2);
  // 2. the reference's loop (:48-141)
  vnAssignedF = std::vector<int>(vKeysUnF.size(), -1);
  std::vector<bool> occupied = vbOccupiedF;
  int nmatches = 0;
  size_t cur = 0;
  for (size_t iMP = 0; iMP < vPoints.size(); iMP++) {
    const ProjectedPoint& mp = vPoints[iMP];
    if (!mp.inView || cand[iMP].empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (size_t idx : cand[iMP]) {
      const int d = dist[cur++];
      if (occupied[idx]) continue;
      if (vuRightF[idx] > 0) {
        const float er = std::fabs(mp.projXR - vuRightF[idx]);
        if (er > radius[iMP]) continue;
      }
      if (d < bestDist) { bestDist2 = bestDist; bestDist = d; bestLevel2 = bestLevel; bestLevel = vKeysUnF[idx].octave; bestIdx = static_cast<int>(idx); }
      else if (d < bestDist2) { bestLevel2 = vKeysUnF[idx].octave; bestDist2 = d; }
    }
    if (bestDist <= TH_HIGH) {
      if (bestLevel == bestLevel2 && bestDist > mfNNratio * bestDist2) continue;
      if (bestLevel != bestLevel2 || bestDist <= mfNNratio * bestDist2) {
        vnAssignedF[bestIdx] = static_cast<int>(iMP);
        occupied[bestIdx] = mp.hasObservations;
        nmatches++;
      }
    }
  }
  return nmatches;
}

int XFBmatcher::SearchByProjection(const std::vector<LastFramePoint>& vLast, const cv::Mat& descLast, const std::vector<cv::KeyPoint>& vKeysUnCur,
                                   const cv::Mat& descCur, const std::vector<bool>& vbOccupiedCur, const std::vector<float>& vuRightCur, float minX,
                                   float minY, float maxX, float maxY, float scaleFactor, float th, float mbf, bool bForward, bool bBackward,
                                   std::vector<int>& vnAssignedCur) const {
  const Grid grid(vKeysUnCur, minX, minY, maxX, maxY);
  std::vector<std::vector<size_t> > cand(vLast.size());
  std::vector<float> radius(vLast.size(), 0.f);
  std::vector<int32_t> p1, p2;
  for (size_t i = 0; i < vLast.size(); i++) {
    const LastFramePoint& lp = vLast[i];
    if (!lp.valid) continue;
    const int nLastOctave = lp.octave;
    float sf = 1.0f;                                 // CurrentFrame.mvScaleFactors[nLastOctave]
    for (int l = 0; l < nLastOctave; ++l) sf *= scaleFactor;
    radius[i] = th * sf;
    if (bForward) cand[i] = grid.area(vKeysUnCur, lp.u, lp.v, radius[i], nLastOctave);
    else if (bBackward) cand[i] = grid.area(vKeysUnCur, lp.u, lp.v, radius[i], 0, nLastOctave);
    else cand[i] = grid.area(vKeysUnCur, lp.u, lp.v, radius[i], nLastOctave - 1, nLastOctave + 1);
    for (size_t i2 : cand[i]) { p1.push_back(static_cast<int32_t>(i)); p2.push_back(static_cast<int32_t>(i2)); }
  }
  const std::vector<int32_t> dist = PairDistances(descLast, descCur, p1, p2);
  vnAssignedCur = std::vector<int>(vKeysUnCur.size(), -1);
  std::vector<bool> occupied = vbOccupiedCur;
  int nmatches = 0;
  size_t cur = 0;
  for (size_t i = 0; i < vLast.size(); i++) {
    const LastFramePoint& lp = vLast[i];
    if (!lp.valid || cand[i].empty()) continue;
    int bestDist = 256, bestIdx2 = -1;
    for (size_t i2 : cand[i]) {
      const int d = dist[cur++];
      if (occupied[i2]) continue;
      if (vuRightCur[i2] > 0) {
        const float ur = lp.u - mbf * lp.invzc;
        const float er = std::fabs(ur - vuRightCur[i2]);
        if (er > radius[i]) continue;
      }
      if (d < bestDist) { bestDist = d; bestIdx2 = static_cast<int>(i2); }
    }
    // (reference :1955-1958 tests only bestDist <= TH_HIGH; with TH_HIGH = 1000 > the initial 256 it then writes mvpMapPoints[-1]
    //  when no candidate was taken -- undefined behaviour there, "no match" here)
    if (bestDist <= TH_HIGH && bestIdx2 >= 0) {
      vnAssignedCur[bestIdx2] = static_cast<int>(i);
      occupied[bestIdx2] = lp.hasObservations;
      nmatches++;
      // rotation histogram (:1967-1985, :2047-2069): every XFeat keypoint has angle -1 => one bin, nothing is removed
    }
  }
  return nmatches;
}

// The window search shared by the loop-closing (:612-717) and relocalisation (:2074-2190) SearchByProjection overloads: candidates in
// the window whose octave lies in [predictedLevel - 1, predictedLevel + levelHiOffset], not taken yet; the closest one is taken when
// its distance is <= threshold.
int XFBmatcher::WindowAssign(const std::vector<WindowQuery>& vQueries, const cv::Mat& descMP, const std::vector<cv::KeyPoint>& vKeys,
                             const cv::Mat& descKF, const std::vector<bool>& vbTaken, float minX, float minY, float maxX, float maxY,
                             int levelHiOffset, float threshold, std::vector<int>& vnAssigned) const {
  const Grid grid(vKeys, minX, minY, maxX, maxY);
  std::vector<std::vector<size_t> > cand(vQueries.size());
  std::vector<int32_t> p1, p2;
  for (size_t iMP = 0; iMP < vQueries.size(); iMP++) {
    const WindowQuery& q = vQueries[iMP];
    if (!q.valid) continue;
    for (size_t idx : grid.area(vKeys, q.u, q.v, q.radius)) {
      const int kpLevel = vKeys[idx].octave;
      if (kpLevel < q.predictedLevel - 1 || kpLevel > q.predictedLevel + levelHiOffset) continue;   // static filter, applied before listing
      cand[iMP].push_back(idx);
      p1.push_back(static_cast<int32_t>(iMP)); p2.push_back(static_cast<int32_t>(idx));
    }
  }
  const std::vector<int32_t> dist = PairDistances(descMP, descKF, p1, p2);
  vnAssigned = std::vector<int>(vKeys.size(), -1);
  std::vector<bool> taken = vbTaken;
  int nmatches = 0;
  size_t cur = 0;
  for (size_t iMP = 0; iMP < vQueries.size(); iMP++) {
    if (!vQueries[iMP].valid || cand[iMP].empty()) continue;
    int bestDist = 256, bestIdx = -1;
    for (size_t idx : cand[iMP]) {
      const int d = dist[cur++];
      if (taken[idx]) continue;
      if (d < bestDist) { bestDist = d; bestIdx = static_cast<int>(idx); }
    }
    if (bestDist <= threshold && bestIdx >= 0) {      // (bestIdx >= 0 only matters for thresholds >= 256, where the reference indexes [-1])
      vnAssigned[bestIdx] = static_cast<int>(iMP);
      taken[bestIdx] = true;
      nmatches++;
    }
  }
  return nmatches;
}

int XFBmatcher::SearchByProjection(const std::vector<WindowQuery>& vQueries, const cv::Mat& descMP, const std::vector<cv::KeyPoint>& vKeysUnKF,
                                   const cv::Mat& descKF, const std::vector<bool>& vbMatchedKF, float minX, float minY, float maxX, float maxY,
                                   float ratioHamming, std::vector<int>& vnAssignedKF) const {
  // GetFeaturesInArea(u, v, radius) + `kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel` (:668, :688-691); `bestDist <= TH_LOW * ratioHamming` (:705)
  return WindowAssign(vQueries, descMP, vKeysUnKF, descKF, vbMatchedKF, minX, minY, maxX, maxY, 0, TH_LOW * ratioHamming, vnAssignedKF);
}

int XFBmatcher::SearchByProjectionReloc(const std::vector<WindowQuery>& vQueries, const cv::Mat& descMP, const std::vector<cv::KeyPoint>& vKeysUnCur,
                                        const cv::Mat& descCur, const std::vector<bool>& vbOccupiedCur, float minX, float minY, float maxX, float maxY,
                                        int ORBdist, std::vector<int>& vnAssignedCur) const {
  // GetFeaturesInArea(u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1) (:2124); `bestDist <= ORBdist` (:2147)
  return WindowAssign(vQueries, descMP, vKeysUnCur, descCur, vbOccupiedCur, minX, minY, maxX, maxY, 1, static_cast<float>(ORBdist), vnAssignedCur);
}

void XFBmatcher::FuseSearch(const std::vector<WindowQuery>& vQueries, const cv::Mat& descMP, const std::vector<cv::KeyPoint>& vKeysUnKF,
                            const std::vector<float>& vuRightKF, const std::vector<float>& vInvLevelSigma2, const cv::Mat& descKF, float minX,
                            float minY, float maxX, float maxY, bool bChi2, std::vector<int>& vnBestIdx, std::vector<int>& vnBestDist, int initDist) const {
  const Grid grid(vKeysUnKF, minX, minY, maxX, maxY);
  std::vector<std::vector<size_t> > cand(vQueries.size());
  std::vector<int32_t> p1, p2;
  for (size_t i = 0; i < vQueries.size(); i++) {
    const WindowQuery& q = vQueries[i];
    if (!q.valid) continue;
    for (size_t idx : grid.area(vKeysUnKF, q.u, q.v, q.radius)) {
      const cv::KeyPoint& kp = vKeysUnKF[idx];
      const int kpLevel = kp.octave;
      if (kpLevel < q.predictedLevel - 1 || kpLevel > q.predictedLevel) continue;
      if (bChi2) {
        if (vuRightKF[idx] >= 0) {   // reprojection error in stereo (:1438-1449)
          const float ex = q.u - kp.pt.x, ey = q.v - kp.pt.y, er = q.ur - vuRightKF[idx];
          const float e2 = ex * ex + ey * ey + er * er;
          if (e2 * vInvLevelSigma2[kpLevel] > 7.8) continue;
        } else {                      // (:1451-1459)
          const float ex = q.u - kp.pt.x, ey = q.v - kp.pt.y;
          const float e2 = ex * ex + ey * ey;
          if (e2 * vInvLevelSigma2[kpLevel] > 5.99) continue;
        }
      }
      cand[i].push_back(idx);
      p1.push_back(static_cast<int32_t>(i)); p2.push_back(static_cast<int32_t>(idx));
    }
  }
  const std::vector<int32_t> dist = PairDistances(descMP, descKF, p1, p2);
  vnBestIdx.assign(vQueries.size(), -1);
  vnBestDist.assign(vQueries.size(), initDist);
  size_t cur = 0;
  for (size_t i = 0; i < vQueries.size(); i++) {
    int bestDist = initDist, bestIdx = -1;
    for (size_t idx : cand[i]) {
      const int d = dist[cur++];
      if (d < bestDist) { bestDist = d; bestIdx = static_cast<int>(idx); }
    }
    vnBestIdx[i] = bestIdx; vnBestDist[i] = bestDist;
  }
}

int XFBmatcher::SearchBySim3(const std::vector<WindowQuery>& vQueries1, const cv::Mat& descMP1, const std::vector<cv::KeyPoint>& vKeysUn1,
                             const cv::Mat& desc1, const std::vector<WindowQuery>& vQueries2, const cv::Mat& descMP2,
                             const std::vector<cv::KeyPoint>& vKeysUn2, const cv::Mat& desc2, float minX, float minY, float maxX, float maxY,
                             std::vector<int>& vnMatches12) const {
  const std::vector<float> none;
  std::vector<int> bi1, bd1, bi2, bd2;
  // transform from KF1 to KF2 and search (:1690-1760), then from KF2 to KF1 (:1762-1833): bestDist starts at INT_MAX, no chi-square test
  FuseSearch(vQueries1, descMP1, vKeysUn2, none, none, desc2, minX, minY, maxX, maxY, false, bi1, bd1, INT_MAX);
  FuseSearch(vQueries2, descMP2, vKeysUn1, none, none, desc1, minX, minY, maxX, maxY, false, bi2, bd2, INT_MAX);
  // check agreement (:1835-1850)
  vnMatches12.assign(vQueries1.size(), -1);
  int nFound = 0;
  for (size_t i1 = 0; i1 < vQueries1.size(); i1++) {
    const int idx2 = (bd1[i1] <= TH_HIGH) ? bi1[i1] : -1;
    if (idx2 >= 0) {
      const int idx1 = (bd2[idx2] <= TH_HIGH) ? bi2[idx2] : -1;
      if (idx1 == static_cast<int>(i1)) { vnMatches12[i1] = idx2; nFound++; }
    }
  }
  return nFound;
}

std::vector<int> XFBmatcher::ComputeDistinctiveDescriptors(const cv::Mat& desc, const std::vector<int>& offsets) const {
  const size_t nsets = offsets.empty() ? 0 : offsets.size() - 1;
  std::vector<int32_t> p1, p2;
  for (size_t s = 0; s < nsets; ++s)
    for (int i = offsets[s]; i < offsets[s + 1]; ++i)
      for (int j = i + 1; j < offsets[s + 1]; ++j) { p1.push_back(i); p2.push_back(j); }
  const std::vector<int32_t> dist = PairDistances(desc, desc, p1, p2);
  std::vector<int> best(nsets, -1);
  size_t cur = 0;
  for (size_t s = 0; s < nsets; ++s) {
    const int N = offsets[s + 1] - offsets[s];
    if (N <= 0) continue;
    std::vector<int> Distances(static_cast<size_t>(N) * N, 0);
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) { const int d = dist[cur++]; Distances[i * N + j] = d; Distances[j * N + i] = d; }
    int BestMedian = INT_MAX, BestIdx = 0;
    for (int i = 0; i < N; ++i) {
      std::vector<int> vDists(Distances.begin() + i * N, Distances.begin() + (i + 1) * N);
      std::sort(vDists.begin(), vDists.end());
      const int median = vDists[static_cast<size_t>(0.5 * (N - 1))];
      if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    best[s] = BestIdx;
  }
  return best;
}

}  // namespace ORB_SLAM3
