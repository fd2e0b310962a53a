// XFextractor.cc -- host side of the drop-in: replaces the reference's src/XFextractor.cc.
// Everything the reference did in libtorch (src/XFextractor.cc:258-316) now happens inside
// xfb_extract(); what stays on the host is what was host code in the reference too: the scale
// pyramid bookkeeping of the constructor (:80-112) and the keypoint / descriptor packing loop with
// its mono / lapping-area placement (:310-356).
#include "XFextractor.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>

namespace ORB_SLAM3 {

std::string XFextractor::DefaultWeightsPath() {
  if (const char* env = std::getenv("XFEAT_B200_WEIGHTS")) return env;
  // like the reference (src/XFextractor.cc:151-159) the default is relative to this source file
  std::string here = __FILE__;
  const size_t cut = here.find_last_of('/');
  here = (cut == std::string::npos) ? std::string(".") : here.substr(0, cut);
  return here + "/../weights/xfeat_b200.bin";
}

XFextractor::XFextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
  // scale pyramid tables consumed by Frame (src/Frame.cc:299-305); values as src/XFextractor.cc:80-96
  mvScaleFactor.assign(nlevels, 1.0f);
  mvLevelSigma2.assign(nlevels, 1.0f);
  for (int i = 1; i < nlevels; i++) {
    mvScaleFactor[i] = static_cast<float>(mvScaleFactor[i - 1] * scaleFactor);
    mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
  }
  mvInvScaleFactor.resize(nlevels);
  mvInvLevelSigma2.resize(nlevels);
  for (int i = 0; i < nlevels; i++) {
    mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
    mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
  }
  mvImagePyramid.resize(nlevels);
  // per-level feature budget (:100-112); unused by XFeat, kept for interface parity
  mnFeaturesPerLevel.resize(nlevels);
  const float factor = 1.0f / static_cast<float>(scaleFactor);
  float per = nfeatures * (1 - factor) / (1 - static_cast<float>(std::pow(static_cast<double>(factor), static_cast<double>(nlevels))));
  int sum = 0;
  for (int level = 0; level < nlevels - 1; level++) {
    mnFeaturesPerLevel[level] = cvRound(per);
    sum += mnFeaturesPerLevel[level];
    per *= factor;
  }
  if (nlevels > 0) mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);

  if (nfeatures < 1 || nfeatures > 8192) throw std::invalid_argument("XFextractor: nfeatures must be in [1, 8192]");
  const std::string path = DefaultWeightsPath();
  std::ifstream is(path, std::ios::binary);
  if (!is) throw std::runtime_error("XFextractor: cannot open weight blob " + path);
  weights_.assign(std::istreambuf_iterator<char>(is), std::istreambuf_iterator<char>());
  std::cout << "XFeat model weights loaded successfully!" << std::endl;   // same banner as :138
  std::cout << "Device: cuda (libxfeat_b200, sm_100a)" << std::endl;
}

XFextractor::~XFextractor() { xfb_destroy(ctx_); }

void XFextractor::EnsureContext(int h, int w) {
  if (ctx_ && h <= ctx_h_ && w <= ctx_w_) return;
  // grow only: keep the capacity of every size seen so far, so that alternating sizes (480x752, then 512x640) re-create once
  h = std::max(h, ctx_h_);
  w = std::max(w, ctx_w_);
  xfb_destroy(ctx_);
  ctx_ = nullptr;
  int device = 0;
  if (const char* env = std::getenv("XFEAT_B200_DEVICE")) device = std::atoi(env);
  const int rc = xfb_create(&ctx_, weights_.data(), weights_.size(), device, h, w, 1, nfeatures);
  if (rc != XFB_OK) throw std::runtime_error(std::string("XFextractor: xfb_create failed: ") + xfb_last_error(nullptr));
  ctx_h_ = h; ctx_w_ = w;
  kpt_xy_.resize(static_cast<size_t>(nfeatures) * 2);
  score_.resize(nfeatures);
  desc_.resize(static_cast<size_t>(nfeatures) * XFB_DESC_DIM);
}

int XFextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                            cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
  if (_image.empty()) return -1;                                   // :253-254
  cv::Mat image = _image.getMat();
  assert(image.type() == CV_8UC1);                                 // :257
  if (image.channels() != 1 && image.channels() != 3)
    throw std::invalid_argument("Unsupported number of channels in the input image.");   // :179
  EnsureContext(image.rows, image.cols);

  int32_t n_valid = 0;
  const int rc = xfb_extract(ctx_, image.data, image.rows, image.cols, static_cast<int>(image.step), nfeatures, 0.05f, &n_valid,
                             kpt_xy_.data(), score_.data(), desc_.data());
  if (rc != XFB_OK) throw std::runtime_error(std::string("XFextractor: xfb_extract failed: ") + xfb_last_error(ctx_));

  // ---- packing, as src/XFextractor.cc:310-356 ---------------------------------------------------
  _keypoints = std::vector<cv::KeyPoint>(nfeatures);
  cv::Mat desc_mat(cv::Size(64, nfeatures), CV_32F, cv::Scalar(0));
  int monoIndex = 0, stereoIndex = nfeatures - 1;
  for (int i = 0; i < n_valid; i++) {
    const float x = kpt_xy_[2 * i], y = kpt_xy_[2 * i + 1];
    cv::KeyPoint keypoint(x, y, 1, -1, score_[i]);
    int dst;
    if (x >= vLappingArea[0] && x <= vLappingArea[1]) dst = stereoIndex--;
    else dst = monoIndex++;
    _keypoints.at(dst) = keypoint;
    std::memcpy(desc_mat.ptr<float>(dst), desc_.data() + static_cast<size_t>(i) * 64, 64 * sizeof(float));
  }
  if (n_valid > 0) desc_mat.rowRange(cv::Range(0, static_cast<int>(_keypoints.size()))).copyTo(_descriptors);
  else _descriptors.release();
  return monoIndex;
}

}  // namespace ORB_SLAM3
