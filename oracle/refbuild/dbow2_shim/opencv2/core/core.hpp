// TEST INFRASTRUCTURE ONLY -- the few cv:: names the reference's vendored DBoW2 (thirdparty/DBoW2/DBoW2/FORB.cpp,
// TemplatedVocabulary.h) touches, so that those files compile UNCHANGED without OpenCV (absent from this image):
// cv::Mat as a row-major byte buffer with create / zeros / ptr<T> / clone / rows / cols (FORB.cpp:46-63,:87-88,:108,:122-123),
// and a cv::FileStorage / cv::FileNode pair that only has to COMPILE (TemplatedVocabulary.h:1456-1625: the YAML save / load
// are virtual, hence instantiated, but never called here -- the reference itself loads ORBvoc.txt through loadFromTextFile).
#pragma once
#include <cmath>
#include <math.h>
#include <sstream>
#include <iostream>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5

namespace cv {

class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  // a 1 x cols view of externally owned memory (what Converter::toDescriptorVector's Descriptors.row(j) is)
  Mat(int r, int c, int type, void* data) : rows(r), cols(c), type_(type), view_(static_cast<unsigned char*>(data)) {}
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; view_ = nullptr;
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * elem(), 0);
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    if (rows * cols) std::memcpy(m.ptr<unsigned char>(), ptr<unsigned char>(), (size_t)rows * cols * elem());
    return m;
  }
  void release() { rows = cols = 0; view_ = nullptr; buf_.reset(); }
  bool empty() const { return rows * cols == 0; }
  int type() const { return type_; }
  template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(base() + (size_t)r * cols * elem()); }
  template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(base() + (size_t)r * cols * elem()); }

 private:
  size_t elem() const { return type_ == CV_32F ? 4 : 1; }
  unsigned char* base() const { return view_ ? view_ : (buf_ ? buf_->data() : nullptr); }
  int type_ = CV_8U;
  unsigned char* view_ = nullptr;
  std::shared_ptr<std::vector<unsigned char>> buf_;
};

class FileNode {
 public:
  FileNode operator[](const std::string&) const { std::abort(); }
  FileNode operator[](const char*) const { std::abort(); }
  FileNode operator[](int) const { std::abort(); }
  size_t size() const { std::abort(); }
  operator int() const { std::abort(); }
  operator double() const { std::abort(); }
  operator std::string() const { std::abort(); }
};

class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage(const char*, int) {}
  bool isOpened() const { return false; }
  FileNode operator[](const std::string&) const { std::abort(); }
};
template <class T> FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

}  // namespace cv
