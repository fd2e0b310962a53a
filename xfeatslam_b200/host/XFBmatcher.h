// XFBmatcher.h -- host-side companion of the reference's ORBmatcher for XFeat descriptors.
//
// The reference's matchers all reduce to ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2242-2250)
// called pair by pair from host loops (18 call sites).  The drop-in keeps those host
// loops (they walk Frames / KeyFrames / MapPoints under the callers' mutexes) and replaces the
// distance evaluations by ONE all-pairs GPU computation per frame pair:
//   * DistanceTable   : D(i, j) == ORBmatcher::DescriptorDistance(desc1.row(i), desc2.row(j)), bit-exact
//   * match()         : the definition for the declared-but-undefined ORBmatcher::match
//                       (include/ORBmatcher.h:77; spec = the commented-out body, src/ORBmatcher.cc:340-406):
//                       mutual nearest neighbours, distance = sqrt(2 (1 - cos)) = ||a - b||
//   * SearchForInitialization : src/ORBmatcher.cc:833-948 replayed over a DistanceTable, with
//                       Frame::GetFeaturesInArea (src/Frame.cc:850-916) restated on plain arrays.
// See INTEGRATION.md for the 3-line patch that routes ORBmatcher::DescriptorDistance through a table.
#ifndef XFBMATCHER_H
#define XFBMATCHER_H

#include <vector>

#include <opencv2/opencv.hpp>

#include "xfeat_b200.h"

namespace ORB_SLAM3 {

class XFBmatcher {
 public:
  static const int TH_HIGH = 1000;   // src/ORBmatcher.cc:34 (USE_ORB unset)
  static const int TH_LOW = 100;     // :35
  static const int HISTO_LENGTH = 30;

  XFBmatcher(xfb_ctx* ctx, float nnratio = 0.6f, bool checkOri = true);

  // All-pairs DescriptorDistance, computed once on the GPU.
  class DistanceTable {
   public:
    int operator()(int i1, int i2) const { return d_[static_cast<size_t>(i1) * n2_ + i2]; }
    int rows() const { return n1_; }
    int cols() const { return n2_; }
   private:
    friend class XFBmatcher;
    std::vector<int32_t> d_;
    int n1_ = 0, n2_ = 0;
  };
  DistanceTable ComputeDistances(const cv::Mat& desc1, const cv::Mat& desc2) const;

  // Brute-force mutual-NN matcher (ORBmatcher::match slot).
  void match(cv::Mat _frame1_desc, cv::Mat _frame2_desc, std::vector<cv::DMatch>& _matches) const;

  // ORBmatcher::SearchForInitialization on plain containers.  (minX, minY, maxX, maxY) are the
  // Frame's undistorted image bounds (mnMinX ... of src/Frame.cc), keys*Un the undistorted keypoints.
  int SearchForInitialization(const std::vector<cv::KeyPoint>& keys1Un, const cv::Mat& desc1, const std::vector<cv::KeyPoint>& keys2Un,
                              const cv::Mat& desc2, float minX, float minY, float maxX, float maxY,
                              std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10) const;

 private:
  xfb_ctx* ctx_;
  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM3
#endif
