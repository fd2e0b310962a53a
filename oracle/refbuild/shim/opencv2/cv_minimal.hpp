// cv_minimal.hpp -- TEST INFRASTRUCTURE ONLY (not product code).
//
// A tiny stand-in for the handful of OpenCV core types that the reference's
// XFextractor.cc / XFeat.cc touch (cv::Mat as a dense row-major buffer,
// cv::KeyPoint, cv::InputArray / cv::OutputArray proxies, Size / Scalar /
// Range / Point2f, cvRound/cvFloor/cvCeil).  OpenCV's C++ headers are not in
// this image; this header lets (a) the reference sources compile *unchanged*
// into oracle/_ref and (b) our own drop-in host class
// (xfeatslam_b200/host/XFextractor.cc) be compile- and parity-tested here.
// A real xfeatSLAM build uses the real <opencv2/...> headers instead.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

static inline int cvRound(double v) { return static_cast<int>(std::lrint(v)); }
static inline int cvFloor(double v) { return static_cast<int>(std::floor(v)); }
static inline int cvCeil(double v) { return static_cast<int>(std::ceil(v)); }

namespace cv {

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};

struct Scalar {
  double val[4];
  Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) {
    val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3;
  }
};

struct Range {
  int start, end;
  Range(int s, int e) : start(s), end(e) {}
};

struct Point2f {
  float x, y;
  Point2f() : x(0.f), y(0.f) {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct KeyPoint {
  Point2f pt;
  float size;
  float angle;
  float response;
  int octave;
  int class_id;
  KeyPoint() : pt(), size(0.f), angle(-1.f), response(0.f), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1.f, float response_ = 0.f,
           int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_),
        class_id(class_id_) {}
};

struct DMatch {
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4e38f) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

class _OutputArray;

// Dense 2-D matrix with shared ownership of its storage (row views alias the
// parent like cv::Mat headers do).
class Mat {
 public:
  uchar* data;
  int rows, cols;
  size_t step;  // bytes per row

  Mat() : data(nullptr), rows(0), cols(0), step(0), type_(0) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size s, int type) : Mat() { create(s.height, s.width, type); }
  Mat(Size s, int type, const Scalar& fill) : Mat() {
    create(s.height, s.width, type);
    setTo(fill);
  }
  // borrow external memory (no ownership), like cv::Mat(rows, cols, type, ptr)
  Mat(int r, int c, int type, void* ext, size_t step_bytes = 0) : Mat() {
    rows = r; cols = c; type_ = type;
    step = step_bytes ? step_bytes : static_cast<size_t>(c) * elemSize();
    data = static_cast<uchar*>(ext);
  }

  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step = static_cast<size_t>(c) * elemSize();
    store_ = std::make_shared<std::vector<uchar>>(static_cast<size_t>(r) * step);
    data = store_->empty() ? nullptr : store_->data();
  }
  void setTo(const Scalar& s) {
    if (!data) return;
    if (depth() == CV_32F) {
      for (int r = 0; r < rows; ++r) {
        float* p = reinterpret_cast<float*>(data + r * step);
        for (int c = 0; c < cols * channels(); ++c) p[c] = static_cast<float>(s.val[0]);
      }
    } else {
      for (int r = 0; r < rows; ++r)
        std::memset(data + r * step, static_cast<int>(s.val[0]), static_cast<size_t>(cols) * elemSize());
    }
  }

  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
  size_t elemSize() const { return (depth() == CV_8U ? 1u : 4u) * static_cast<size_t>(channels()); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }

  template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + r * step); }
  template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + r * step); }
  template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }

  Mat row(int r) const { return rowRange(Range(r, r + 1)); }
  Mat rowRange(const Range& rg) const {
    Mat v;
    v.store_ = store_;
    v.type_ = type_;
    v.rows = rg.end - rg.start;
    v.cols = cols;
    v.step = step;
    v.data = data + static_cast<size_t>(rg.start) * step;
    return v;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; ++r)
      std::memcpy(m.data + r * m.step, data + r * step, static_cast<size_t>(cols) * elemSize());
    return m;
  }
  void release() {
    store_.reset();
    data = nullptr;
    rows = cols = 0;
    step = 0;
  }
  inline void copyTo(const _OutputArray& dst) const;

 private:
  std::shared_ptr<std::vector<uchar>> store_;
  int type_;
};

class _InputArray {
 public:
  _InputArray() : m_(nullptr) {}
  _InputArray(const Mat& m) : m_(&m) {}
  bool empty() const { return m_ == nullptr || m_->empty(); }
  Mat getMat() const { return m_ ? *m_ : Mat(); }

 private:
  const Mat* m_;
};

class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  _OutputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}  // row views are temporaries
  void release() const { m_->release(); }
  Mat& getMatRef() const { return *m_; }

 private:
  Mat* m_;
};

inline void Mat::copyTo(const _OutputArray& dst) const {
  Mat& d = dst.getMatRef();
  if (d.data == nullptr || d.rows != rows || d.cols != cols || d.type() != type_) d.create(rows, cols, type_);
  for (int r = 0; r < rows; ++r)
    std::memcpy(d.data + r * d.step, data + r * step, static_cast<size_t>(cols) * elemSize());
}

typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

}  // namespace cv
