// ref_harness.cc -- TEST INFRASTRUCTURE ONLY (not product code).
//
// Driver around the UNMODIFIED reference sources /root/reference/src/XFeat.cc and
// /root/reference/src/XFextractor.cc (compiled where they lie, see oracle/refbuild/Makefile)
// linked against the pip libtorch (CPU).  It is the executable form of the oracle
// (SURVEY.md section 8c): it pins the Python restatement in oracle/xfeat_oracle.py, produces the
// golden vectors under tests/golden/, and is the "reference" CPU baseline of bench.py.
//
//   ref_xfeat dump  <frame.u8> <H> <W> <nfeatures> <lap0> <lap1> <out.bin> [threads]
//   ref_xfeat bench <frames.u8> <H> <W> <nfeatures> <nframes> <warmup> <threads>
//
// "dump" writes a flat record stream (see write_rec) with every stage of
// XFextractor::operator() (reference file:line in the record comments below).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "XFextractor.h"

namespace {

enum DType : uint32_t { kF32 = 0, kI64 = 1, kU8 = 2, kI32 = 3 };

void write_rec(std::ofstream& os, const std::string& name, torch::Tensor t) {
  t = t.detach().to(torch::kCPU).contiguous();
  uint32_t code;
  if (t.dtype() == torch::kFloat32) code = kF32;
  else if (t.dtype() == torch::kInt64) code = kI64;
  else if (t.dtype() == torch::kUInt8) code = kU8;
  else if (t.dtype() == torch::kInt32) code = kI32;
  else { t = t.to(torch::kFloat32); code = kF32; }
  uint32_t nlen = static_cast<uint32_t>(name.size());
  uint32_t nd = static_cast<uint32_t>(t.dim());
  os.write(reinterpret_cast<const char*>(&nlen), 4);
  os.write(name.data(), nlen);
  os.write(reinterpret_cast<const char*>(&code), 4);
  os.write(reinterpret_cast<const char*>(&nd), 4);
  for (uint32_t i = 0; i < nd; ++i) {
    int64_t d = t.size(i);
    os.write(reinterpret_cast<const char*>(&d), 8);
  }
  os.write(reinterpret_cast<const char*>(t.data_ptr()), static_cast<std::streamsize>(t.nbytes()));
}

std::vector<unsigned char> read_file(const std::string& path, size_t expect) {
  std::ifstream is(path, std::ios::binary);
  if (!is) { std::cerr << "cannot open " << path << std::endl; std::exit(2); }
  std::vector<unsigned char> buf(expect);
  is.read(reinterpret_cast<char*>(buf.data()), static_cast<std::streamsize>(expect));
  if (static_cast<size_t>(is.gcount()) != expect) { std::cerr << "short read " << path << std::endl; std::exit(2); }
  return buf;
}

// Gives the harness access to the protected stage methods so every intermediate can be dumped by
// calling the reference's own code (no re-implementation here).
class Probe : public ORB_SLAM3::XFextractor {
 public:
  using ORB_SLAM3::XFextractor::XFextractor;

  void dump(cv::Mat& im, std::ofstream& os, int nfeat) {
    torch::NoGradGuard ng;
    // XFextractor.cc:258-264
    torch::Tensor x = parseInput(im);
    float rh1, rw1;
    std::tie(x, rh1, rw1) = preprocessTensor(x);
    write_rec(os, "x_pre", x);
    const int64_t H1 = x.size(2), W1 = x.size(3);

    // XFeat.cc:147-149
    torch::Tensor xm = x.mean(1, true);
    torch::Tensor xn = model->norm->forward(xm);
    write_rec(os, "xn", xn);

    // XFeat.cc:152-156, layer by layer
    auto run_seq = [&](torch::nn::Sequential& seq, torch::Tensor t, const std::string& tag) {
      int i = 0;
      for (auto& child : seq->children()) {
        if (auto* bl = child->as<ORB_SLAM3::BasicLayerImpl>()) t = bl->forward(t);
        else if (auto* cv2 = child->as<torch::nn::Conv2dImpl>()) t = cv2->forward(t);
        else if (auto* sg = child->as<torch::nn::SigmoidImpl>()) t = sg->forward(t);
        else if (auto* ap = child->as<torch::nn::AvgPool2dImpl>()) t = ap->forward(t);
        else { std::cerr << "unknown child in " << tag << std::endl; std::exit(3); }
        write_rec(os, tag + "." + std::to_string(i), t);
        ++i;
      }
      return t;
    };
    torch::Tensor x1 = run_seq(model->block1, xn, "block1");
    torch::Tensor sk = run_seq(model->skip1, xn, "skip1");
    torch::Tensor x2 = run_seq(model->block2, x1 + sk, "block2");
    torch::Tensor x3 = run_seq(model->block3, x2, "block3");
    torch::Tensor x4 = run_seq(model->block4, x3, "block4");
    torch::Tensor x5 = run_seq(model->block5, x4, "block5");

    // whole-model forward (XFeat.cc:135-173) for the three heads
    torch::Tensor M1, K1, Hm;
    std::tie(M1, K1, Hm) = model->forward(x);
    write_rec(os, "feats", M1);
    write_rec(os, "K1", K1);
    write_rec(os, "H1", Hm);

    // XFextractor.cc:273-282
    torch::Tensor M1n = torch::nn::functional::normalize(M1, torch::nn::functional::NormalizeFuncOptions().dim(1));
    write_rec(os, "M1n", M1n);
    torch::Tensor K1h = getKptsHeatmap(K1);
    write_rec(os, "K1h", K1h);
    torch::Tensor mk = NMS(K1h, 0.05, 5);
    write_rec(os, "nms_kpts", mk);
    auto sc_near = nearest->forward(K1h, mk, H1, W1);
    auto sc_bil = bilinear->forward(Hm, mk, H1, W1);
    write_rec(os, "score_nearest", sc_near);
    write_rec(os, "score_bilinear", sc_bil);
    auto scores = (sc_near * sc_bil).squeeze(-1);
    auto mask = torch::all(mk == 0, -1);
    scores.masked_fill_(mask, -1);
    write_rec(os, "scores_all", scores);
    // descriptors at *all* NMS keypoints (order independent check of XFextractor.cc:298-301)
    torch::Tensor fall = bilinear->forward(M1n, mk, H1, W1);
    fall = torch::nn::functional::normalize(fall, torch::nn::functional::NormalizeFuncOptions().dim(-1));
    write_rec(os, "desc_all", fall);
    (void)nfeat; (void)rh1; (void)rw1;
  }
};

void pack_outputs(const std::vector<cv::KeyPoint>& kps, const cv::Mat& desc, int ret, std::ofstream& os) {
  const int n = static_cast<int>(kps.size());
  torch::Tensor k = torch::zeros({n, 7}, torch::kFloat32);
  auto ka = k.accessor<float, 2>();
  for (int i = 0; i < n; ++i) {
    ka[i][0] = kps[i].pt.x; ka[i][1] = kps[i].pt.y; ka[i][2] = kps[i].response;
    ka[i][3] = kps[i].size; ka[i][4] = kps[i].angle;
    ka[i][5] = static_cast<float>(kps[i].octave); ka[i][6] = static_cast<float>(kps[i].class_id);
  }
  write_rec(os, "out_keypoints", k);
  torch::Tensor d = torch::zeros({desc.rows, desc.cols}, torch::kFloat32);
  for (int r = 0; r < desc.rows; ++r)
    std::memcpy(d.data_ptr<float>() + static_cast<size_t>(r) * desc.cols, desc.ptr<float>(r), sizeof(float) * desc.cols);
  write_rec(os, "out_descriptors", d);
  write_rec(os, "out_ret", torch::tensor({static_cast<int64_t>(ret)}, torch::kInt64));
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) { std::cerr << "usage: ref_xfeat dump|bench ..." << std::endl; return 1; }
  const std::string mode = argv[1];
  if (mode == "dump") {
    if (argc < 9) { std::cerr << "dump <frame.u8> H W nfeatures lap0 lap1 out.bin [threads]" << std::endl; return 1; }
    const std::string fpath = argv[2];
    const int H = std::atoi(argv[3]), W = std::atoi(argv[4]), nfeat = std::atoi(argv[5]);
    std::vector<int> lap = {std::atoi(argv[6]), std::atoi(argv[7])};
    const std::string opath = argv[8];
    if (argc > 9) torch::set_num_threads(std::atoi(argv[9]));
    auto buf = read_file(fpath, static_cast<size_t>(H) * W);
    cv::Mat im(H, W, CV_8UC1);
    std::memcpy(im.data, buf.data(), buf.size());
    Probe ex(nfeat, 1.2f, 8, 20, 7);
    std::ofstream os(opath, std::ios::binary);
    ex.dump(im, os, nfeat);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
    int ret = ex(im, cv::Mat(), kps, desc, lap);  // the reference entry point, XFextractor.cc:250
    pack_outputs(kps, desc, ret, os);
    std::cerr << "dump ok: ret=" << ret << " nkp=" << kps.size() << " desc=" << desc.rows << "x" << desc.cols << std::endl;
    return 0;
  }
  if (mode == "bench") {
    if (argc < 9) { std::cerr << "bench <frames.u8> H W nfeatures nframes warmup threads" << std::endl; return 1; }
    const std::string fpath = argv[2];
    const int H = std::atoi(argv[3]), W = std::atoi(argv[4]), nfeat = std::atoi(argv[5]);
    const int nframes = std::atoi(argv[6]), warm = std::atoi(argv[7]), threads = std::atoi(argv[8]);
    if (threads > 0) torch::set_num_threads(threads);
    const size_t fsz = static_cast<size_t>(H) * W;
    auto buf = read_file(fpath, fsz * nframes);
    ORB_SLAM3::XFextractor ex(nfeat, 1.2f, 8, 20, 7);
    std::vector<int> lap = {0, 0};
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
    cv::Mat im(H, W, CV_8UC1);
    long total_valid = 0;
    for (int i = 0; i < warm; ++i) {
      std::memcpy(im.data, buf.data() + fsz * (i % nframes), fsz);
      ex(im, cv::Mat(), kps, desc, lap);
    }
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < nframes; ++i) {
      std::memcpy(im.data, buf.data() + fsz * i, fsz);
      int r = ex(im, cv::Mat(), kps, desc, lap);
      total_valid += r;
    }
    auto t1 = std::chrono::steady_clock::now();
    double sec = std::chrono::duration<double>(t1 - t0).count();
    std::printf("{\"frames\": %d, \"seconds\": %.6f, \"fps\": %.4f, \"threads\": %d, \"mono_index_sum\": %ld}\n", nframes, sec,
                nframes / sec, torch::get_num_threads(), total_valid);
    return 0;
  }
  std::cerr << "unknown mode " << mode << std::endl;
  return 1;
}
