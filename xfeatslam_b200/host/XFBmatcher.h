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
//   * SearchByBoW (both overloads), SearchForTriangulation, SearchByProjection, ComputeDistinctiveDescriptors:
//                       the candidate pairs are listed in the reference's visiting order, all their distances come from ONE
//                       xfb_distance_pairs launch, and the reference's sequential accept / reject logic is replayed over them.
//                       MapPoint* / KeyFrame* / Frame& arguments become plain containers (those classes need Eigen / Sophus /
//                       DBoW2); DBoW2::FeatureVector is kept as the std::map it is.
// See INTEGRATION.md for the 3-line patch that routes ORBmatcher::DescriptorDistance through a table.
#ifndef XFBMATCHER_H
#define XFBMATCHER_H

#include <map>
#include <utility>
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

  // DBoW2::FeatureVector (thirdparty/DBoW2/DBoW2/FeatureVector.h): node id -> feature indices, ascending
  typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;

  // ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches), src/ORBmatcher.cc:408-610
  // (F.Nleft == -1).  vbGoodMapPointKF[i] = (vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad()).
  // vnMatchesF[j] = index of the KF feature whose MapPoint frame feature j receives, -1 = NULL.
  int SearchByBoW(const cv::Mat& descKF, const FeatureVector& vFeatVecKF, const std::vector<bool>& vbGoodMapPointKF, const cv::Mat& descF,
                  const FeatureVector& vFeatVecF, std::vector<int>& vnMatchesF) const;
  // ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12), src/ORBmatcher.cc:950-1090.
  // vnMatches12[i1] = index of the KF2 feature whose MapPoint is matched to feature i1 of KF1, -1 = NULL.
  int SearchByBoW(const cv::Mat& desc1, const FeatureVector& vFeatVec1, const std::vector<bool>& vbGoodMapPoint1, const cv::Mat& desc2,
                  const FeatureVector& vFeatVec2, const std::vector<bool>& vbGoodMapPoint2, std::vector<int>& vnMatches12) const;
  // ORBmatcher::SearchForTriangulation, src/ORBmatcher.cc:1092-1331, one pinhole camera per keyframe.
  // vbHasMapPoint = (GetMapPoint(idx) != NULL); vbStereo = (mvuRight[idx] >= 0); F12 = the row-major fundamental matrix of
  // Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:109-112); ep = epipole of camera 1 in image 2 (:1105);
  // sigma2Level0 = mvLevelSigma2[0], scaleFactor0 = mvScaleFactors[0].
  int SearchForTriangulation(const cv::Mat& desc1, const FeatureVector& vFeatVec1, const std::vector<bool>& vbHasMapPoint1,
                             const std::vector<bool>& vbStereo1, const std::vector<cv::KeyPoint>& vKeysUn1, const cv::Mat& desc2,
                             const FeatureVector& vFeatVec2, const std::vector<bool>& vbHasMapPoint2, const std::vector<bool>& vbStereo2,
                             const std::vector<cv::KeyPoint>& vKeysUn2, const float F12[9], const cv::Point2f& ep,
                             std::vector<std::pair<size_t, size_t> >& vMatchedPairs, bool bOnlyStereo, bool bCoarse = false,
                             float sigma2Level0 = 1.0f, float scaleFactor0 = 1.0f) const;

  // One map point of ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)
  struct ProjectedPoint {
    bool inView;            // mbTrackInView && !(bFarPoints && mTrackDepth > thFarPoints) && !isBad()
    float projX, projY;     // mTrackProjX, mTrackProjY
    float projXR;           // mTrackProjXR
    int scaleLevel;         // mnTrackScaleLevel
    float viewCos;          // mTrackViewCos
    bool hasObservations;   // Observations() > 0
  };
  // src/ORBmatcher.cc:42-212 for monocular / RGB-D frames (F.Nleft == -1).  descMP row m = pMP->GetDescriptor();
  // vbOccupiedF[idx] = (F.mvpMapPoints[idx] && F.mvpMapPoints[idx]->Observations() > 0); vuRightF = F.mvuRight.
  // vnAssignedF[idx] = index of the map point written to F.mvpMapPoints[idx], -1 = untouched.
  int SearchByProjection(const std::vector<ProjectedPoint>& vPoints, const cv::Mat& descMP, const std::vector<cv::KeyPoint>& vKeysUnF,
                         const cv::Mat& descF, const std::vector<bool>& vbOccupiedF, const std::vector<float>& vuRightF, float minX, float minY,
                         float maxX, float maxY, float scaleFactor, float th, std::vector<int>& vnAssignedF) const;

  // One feature of LastFrame for ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)
  struct LastFramePoint {
    bool valid;             // pMP && !mvbOutlier[i] && invzc >= 0 && projection inside the image bounds (:1881-1903)
    float u, v;             // CurrentFrame.mpCamera->project(Tcw * pMP->GetWorldPos())
    float invzc;            // 1 / depth in the current camera
    int octave;             // LastFrame.mvKeys[i].octave
    bool hasObservations;   // pMP->Observations() > 0
  };
  // src/ORBmatcher.cc:1861-2072 (TrackWithMotionModel), CurrentFrame.Nleft == -1.  descLast row i = pMP->GetDescriptor() of feature i.
  // bForward / bBackward as computed at :1875-1876.  vnAssignedCur[i2] = last-frame index written to CurrentFrame.mvpMapPoints[i2].
  int SearchByProjection(const std::vector<LastFramePoint>& vLast, const cv::Mat& descLast, const std::vector<cv::KeyPoint>& vKeysUnCur,
                         const cv::Mat& descCur, const std::vector<bool>& vbOccupiedCur, const std::vector<float>& vuRightCur, float minX, float minY,
                         float maxX, float maxY, float scaleFactor, float th, float mbf, bool bForward, bool bBackward,
                         std::vector<int>& vnAssignedCur) const;

  // One projected map point of the loop-closing / fusing searches: everything the reference derives from the pose and the MapPoint
  // before it looks at descriptors (src/ORBmatcher.cc:633-668, :1363-1411)
  struct WindowQuery {
    bool valid;             // passed: not bad / not already found, positive depth, inside the image, distance range, viewing angle
    float u, v;             // pKF->mpCamera->project(Tcw * p3Dw)
    float ur;               // u - bf * invz (Fuse's stereo reprojection test)
    float radius;           // th * pKF->mvScaleFactors[nPredictedLevel]
    int predictedLevel;     // nPredictedLevel: candidates need octave in [predictedLevel - 1, predictedLevel]
  };
  // ORBmatcher::SearchByProjection(KeyFrame*, Sophus::Sim3f&, vpPoints, vpMatched, th, ratioHamming), src/ORBmatcher.cc:612-717
  // (and :719-831, which only adds the vpMatchedKF array).  vbMatchedKF[idx] = (vpMatched[idx] != NULL) on entry;
  // vnAssignedKF[idx] = index of the map point written to vpMatched[idx], -1 = untouched.
  int SearchByProjection(const std::vector<WindowQuery>& vQueries, const cv::Mat& descMP, const std::vector<cv::KeyPoint>& vKeysUnKF,
                         const cv::Mat& descKF, const std::vector<bool>& vbMatchedKF, float minX, float minY, float maxX, float maxY,
                         float ratioHamming, std::vector<int>& vnAssignedKF) const;
  // ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:2074-2190
  // (relocalisation): the same window search with GetFeaturesInArea(u, v, radius, level - 1, level + 1) and `bestDist <= ORBdist`.
  // vbOccupiedCur[i2] = (CurrentFrame.mvpMapPoints[i2] != NULL).
  int SearchByProjectionReloc(const std::vector<WindowQuery>& vQueries, const cv::Mat& descMP, const std::vector<cv::KeyPoint>& vKeysUnCur,
                              const cv::Mat& descCur, const std::vector<bool>& vbOccupiedCur, float minX, float minY, float maxX, float maxY,
                              int ORBdist, std::vector<int>& vnAssignedCur) const;
  // The candidate search of ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) (src/ORBmatcher.cc:1413-1479; the Sim3 overload :1525-1640 has
  // the same search without the chi-square test): closest keypoint per map point that passes the level filter and the reprojection
  // test.  vnBestIdx / vnBestDist: -1 / 256 where nothing qualified; the caller applies `bestDist <= TH_LOW` and Replace / AddObservation.
  void FuseSearch(const std::vector<WindowQuery>& vQueries, const cv::Mat& descMP, const std::vector<cv::KeyPoint>& vKeysUnKF,
                  const std::vector<float>& vuRightKF, const std::vector<float>& vInvLevelSigma2, const cv::Mat& descKF, float minX, float minY,
                  float maxX, float maxY, bool bChi2, std::vector<int>& vnBestIdx, std::vector<int>& vnBestDist, int initDist = 256) const;
  // ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th), src/ORBmatcher.cc:1642-1859.  vQueries1[i1] / vQueries2[i2]: the map point of
  // that feature projected into the OTHER keyframe (valid = good MapPoint, not already matched, positive depth, in image, distance range);
  // descMP1 / descMP2 rows = pMP->GetDescriptor().  vnMatches12[i1] = idx2 where both directions agree (vpMatches12[i1] = vpMapPoints2[idx2]).
  int SearchBySim3(const std::vector<WindowQuery>& vQueries1, const cv::Mat& descMP1, const std::vector<cv::KeyPoint>& vKeysUn1, const cv::Mat& desc1,
                   const std::vector<WindowQuery>& vQueries2, const cv::Mat& descMP2, const std::vector<cv::KeyPoint>& vKeysUn2, const cv::Mat& desc2,
                   float minX, float minY, float maxX, float maxY, std::vector<int>& vnMatches12) const;

  // MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:329-403, for many map points in one GPU launch: the observed
  // descriptors of map point s are the rows offsets[s] .. offsets[s+1]-1 of `desc`; returns the chosen row (relative to the set).
  std::vector<int> ComputeDistinctiveDescriptors(const cv::Mat& desc, const std::vector<int>& offsets) const;

 private:
  int WindowAssign(const std::vector<WindowQuery>& vQueries, const cv::Mat& descMP, const std::vector<cv::KeyPoint>& vKeys, const cv::Mat& descKF,
                   const std::vector<bool>& vbTaken, float minX, float minY, float maxX, float maxY, int levelHiOffset, float threshold,
                   std::vector<int>& vnAssigned) const;
  std::vector<int32_t> PairDistances(const cv::Mat& desc1, const cv::Mat& desc2, const std::vector<int32_t>& i1, const std::vector<int32_t>& i2) const;
  xfb_ctx* ctx_;
  float mfNNratio;
  bool mbCheckOrientation;
};

}  // namespace ORB_SLAM3
#endif
