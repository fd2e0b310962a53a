// XFBmatcher.cc -- see XFBmatcher.h.  Host control flow in C++, distances from libxfeat_b200.so.
#include "XFBmatcher.h"

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
  std::vector<size_t> area(const std::vector<cv::KeyPoint>& keys, float x, float y, float r) const {
    std::vector<size_t> out;
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
          if (keys[j].octave > 0) continue;
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
    const std::vector<size_t> vIndices2 = grid.area(keys2, vbPrevMatched[i1].x, vbPrevMatched[i1].y, static_cast<float>(windowSize));
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

}  // namespace ORB_SLAM3
