// XFextractor.h -- drop-in replacement of the reference's include/XFextractor.h (lines 32-94).
//
// Same namespace, class name, constructor and operator() signature, getters and public
// mvImagePyramid member as the reference, so Frame.cc / Tracking.cc compile against it unchanged
// (they only use: the ctor at src/Tracking.cc:595-598, operator() at src/Frame.cc:611-618 and the
// scale getters at src/Frame.cc:299-305).  The libtorch members are gone: the body calls the C-ABI of
// libxfeat_b200.so (include/xfeat_b200.h).  No CPU fallback -- construction throws std::runtime_error
// when no B200 is available, where the reference would silently run on the CPU.
#ifndef XFEXTRACTOR_H
#define XFEXTRACTOR_H

#include <list>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp>

#include "xfeat_b200.h"

namespace ORB_SLAM3 {

class XFextractor {
 public:
  XFextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
  ~XFextractor();
  XFextractor(const XFextractor&) = delete;
  XFextractor& operator=(const XFextractor&) = delete;

  // Same contract as the reference (src/XFextractor.cc:250-357): returns -1 on an empty image,
  // otherwise monoIndex; _keypoints always has exactly nfeatures entries; mask is ignored.
  int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors,
                 std::vector<int>& vLappingArea);

  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  std::vector<cv::Mat> mvImagePyramid;

  // extras (not in the reference): where the weight blob is looked up, and the live context
  static std::string DefaultWeightsPath();
  xfb_ctx* context() { return ctx_; }

 protected:
  void EnsureContext(int h, int w);

  int nfeatures;
  double scaleFactor;
  int nlevels;
  int iniThFAST;
  int minThFAST;
  std::vector<int> mnFeaturesPerLevel;
  std::vector<float> mvScaleFactor;
  std::vector<float> mvInvScaleFactor;
  std::vector<float> mvLevelSigma2;
  std::vector<float> mvInvLevelSigma2;

  xfb_ctx* ctx_ = nullptr;
  int ctx_h_ = 0, ctx_w_ = 0;
  std::vector<unsigned char> weights_;
  std::vector<float> kpt_xy_, score_, desc_;
};

}  // namespace ORB_SLAM3

#endif
