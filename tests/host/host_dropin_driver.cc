// host_dropin_driver.cc -- exercises the C++ drop-in classes (xfeatslam_b200/host) the way
// Frame::ExtractXF (src/Frame.cc:611-618) and Tracking (src/Tracking.cc:2518) call the reference.
// Built by tests/test_gpu_host_dropin.py against the cv stand-in header; prints nothing but writes
// raw float files the Python test compares with the reference's golden output / the C oracle.
//   extract <frame.u8> H W nfeatures lap0 lap1 <out_prefix>
//   init    <descA.f32> nA <kpA.f32> <descB.f32> nB <kpB.f32> W H window ratio <out.i32>
//   bow      <vocab.txt> <desc.f32> n levelsup <out_prefix>   (XFBvocabulary::loadFromTextFile + transform)
//   searches <bundle.bin> <out.bin>   (SearchByBoW x2, SearchForTriangulation, SearchByProjection, ComputeDistinctiveDescriptors)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "XFBmatcher.h"
#include "XFBvocabulary.h"
#ifndef XFB_CPU_STUB
#include "XFextractor.h"
#endif

template <typename T>
static std::vector<T> slurp(const std::string& p, size_t n) {
  std::vector<T> v(n);
  std::ifstream is(p, std::ios::binary);
  is.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(n * sizeof(T)));
  if (static_cast<size_t>(is.gcount()) != n * sizeof(T)) { std::cerr << "short read " << p << std::endl; std::exit(2); }
  return v;
}
template <typename T>
static void spit(const std::string& p, const std::vector<T>& v) {
  std::ofstream os(p, std::ios::binary);
  os.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
}

// ---- array bundle: [name 16 bytes][dtype 1 byte: f / i / b][pad 7][count int64][data] ... (written by the Python test) ----
struct Bundle {
  std::map<std::string, std::vector<char> > a;
  explicit Bundle(const std::string& path) {
    std::ifstream is(path, std::ios::binary);
    char hdr[32];
    while (is.read(hdr, 32)) {
      const std::string name(hdr, strnlen(hdr, 16));
      long long count; std::memcpy(&count, hdr + 24, 8);
      const size_t bytes = static_cast<size_t>(count) * (hdr[16] == 'b' ? 1 : 4);
      std::vector<char> v(bytes);
      is.read(v.data(), static_cast<std::streamsize>(bytes));
      a[name] = v;
    }
  }
  template <typename T> std::vector<T> get(const std::string& n) const {
    const std::vector<char>& v = a.at(n);
    std::vector<T> o(v.size() / sizeof(T));
    std::memcpy(o.data(), v.data(), v.size());
    return o;
  }
  std::vector<bool> flags(const std::string& n) const { const std::vector<char>& v = a.at(n); return std::vector<bool>(v.begin(), v.end()); }
  float scalar(const std::string& n) const { return get<float>(n)[0]; }
};
static ORB_SLAM3::XFBmatcher::FeatureVector featvec(const std::vector<int>& node) {
  ORB_SLAM3::XFBmatcher::FeatureVector fv;   // FeatureVector::addFeature (thirdparty/DBoW2/DBoW2/FeatureVector.cpp): push_back in feature order
  for (size_t i = 0; i < node.size(); ++i) if (node[i] >= 0) fv[static_cast<unsigned int>(node[i])].push_back(static_cast<unsigned int>(i));
  return fv;
}
static std::vector<cv::KeyPoint> keypoints(const std::vector<float>& xy) {
  std::vector<cv::KeyPoint> k(xy.size() / 2);
  for (size_t i = 0; i < k.size(); ++i) k[i] = cv::KeyPoint(xy[2 * i], xy[2 * i + 1], 1, -1, 1.f);
  return k;
}

// The GPU build owns a real context through XFextractor; the CPU build of this driver (-DXFB_CPU_STUB, linked against
// tests/host/cpu_stub.cc instead of libxfeat_b200.so) checks the HOST logic only and never dereferences the context.
#ifdef XFB_CPU_STUB
struct ContextOwner { xfb_ctx* context() { return reinterpret_cast<xfb_ctx*>(this); } };
#else
struct ContextOwner {
  ORB_SLAM3::XFextractor ex;
  ContextOwner() : ex(64, 1.2f, 8, 20, 7) {   // the context is created by the first extraction
    cv::Mat tiny(32, 64, CV_8UC1);
    std::memset(tiny.data, 7, 32 * 64);
    std::vector<cv::KeyPoint> tk; cv::Mat td; std::vector<int> lap = {0, 0};
    ex(tiny, cv::Mat(), tk, td, lap);
  }
  xfb_ctx* context() { return ex.context(); }
};
#endif

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "";
#ifndef XFB_CPU_STUB
  if (mode == "extract" && argc >= 9) {
    const int H = std::atoi(argv[3]), W = std::atoi(argv[4]), nfeat = std::atoi(argv[5]);
    std::vector<int> lap = {std::atoi(argv[6]), std::atoi(argv[7])};
    auto buf = slurp<unsigned char>(argv[2], static_cast<size_t>(H) * W);
    cv::Mat im(H, W, CV_8UC1, buf.data());
    ORB_SLAM3::XFextractor ex(nfeat, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> kps;
    cv::Mat desc;
    cv::Mat empty;
    if (ex(empty, cv::Mat(), kps, desc, lap) != -1) return 3;          // empty image -> -1 (src/XFextractor.cc:253)
    const int ret = ex(im, cv::Mat(), kps, desc, lap);
    std::vector<float> k(kps.size() * 7), d(static_cast<size_t>(desc.rows) * desc.cols);
    for (size_t i = 0; i < kps.size(); ++i) {
      k[7 * i] = kps[i].pt.x; k[7 * i + 1] = kps[i].pt.y; k[7 * i + 2] = kps[i].response; k[7 * i + 3] = kps[i].size;
      k[7 * i + 4] = kps[i].angle; k[7 * i + 5] = static_cast<float>(kps[i].octave); k[7 * i + 6] = static_cast<float>(kps[i].class_id);
    }
    for (int r = 0; r < desc.rows; ++r) std::memcpy(d.data() + static_cast<size_t>(r) * desc.cols, desc.ptr<float>(r), sizeof(float) * desc.cols);
    const std::string pre = argv[8];
    spit(pre + ".kp", k);
    spit(pre + ".desc", d);
    std::vector<float> meta = {static_cast<float>(ret), static_cast<float>(kps.size()), static_cast<float>(desc.rows), static_cast<float>(ex.GetLevels()),
                               ex.GetScaleFactor(), ex.GetScaleFactors()[7], ex.GetInverseScaleSigmaSquares()[3]};
    spit(pre + ".meta", meta);
    return 0;
  }
#endif
  if (mode == "init" && argc >= 14) {
    const int nA = std::atoi(argv[3]), nB = std::atoi(argv[6]);
    auto dA = slurp<float>(argv[2], static_cast<size_t>(nA) * 64), kA = slurp<float>(argv[4], static_cast<size_t>(nA) * 2);
    auto dB = slurp<float>(argv[5], static_cast<size_t>(nB) * 64), kB = slurp<float>(argv[7], static_cast<size_t>(nB) * 2);
    const int W = std::atoi(argv[8]), H = std::atoi(argv[9]), window = std::atoi(argv[10]);
    const float ratio = static_cast<float>(std::atof(argv[11]));
    cv::Mat A(nA, 64, CV_32F, dA.data()), B(nB, 64, CV_32F, dB.data());
    std::vector<cv::KeyPoint> ka(nA), kb(nB);
    std::vector<cv::Point2f> prev(nA);
    for (int i = 0; i < nA; ++i) { ka[i] = cv::KeyPoint(kA[2 * i], kA[2 * i + 1], 1, -1, 1.f); prev[i] = ka[i].pt; }
    for (int i = 0; i < nB; ++i) kb[i] = cv::KeyPoint(kB[2 * i], kB[2 * i + 1], 1, -1, 1.f);
    ContextOwner ex;
    ORB_SLAM3::XFBmatcher m(ex.context(), ratio, true);
    std::vector<int> m12;
    const int n = m.SearchForInitialization(ka, A, kb, B, 0.f, 0.f, static_cast<float>(W), static_cast<float>(H), prev, m12, window);
    std::vector<cv::DMatch> mm;
    m.match(A, B, mm);
    std::vector<int> out;
    out.push_back(n);
    out.insert(out.end(), m12.begin(), m12.end());
    out.push_back(static_cast<int>(mm.size()));
    for (auto& x : mm) { out.push_back(x.queryIdx); out.push_back(x.trainIdx); }
    spit(argv[12], out);
    std::vector<float> pv(static_cast<size_t>(nA) * 2);
    for (int i = 0; i < nA; ++i) { pv[2 * i] = prev[i].x; pv[2 * i + 1] = prev[i].y; }
    spit(argv[13], pv);
    return 0;
  }
  if (mode == "bow" && argc >= 7) {
    const int n = std::atoi(argv[4]), levelsup = std::atoi(argv[5]);
    auto d = slurp<float>(argv[3], static_cast<size_t>(n) * 64);
    cv::Mat D(n, 64, CV_32F, d.data());
    ContextOwner ex;
    ORB_SLAM3::XFBvocabulary voc = ORB_SLAM3::XFBvocabulary::loadFromTextFile(ex.context(), argv[2]);
    ORB_SLAM3::XFBvocabulary::BowVector v;
    ORB_SLAM3::XFBvocabulary::FeatureVector fv;
    voc.transform(D, v, fv, levelsup);
    std::vector<double> bow;                       // (word id, value) pairs in map order
    for (auto& kv : v) { bow.push_back(static_cast<double>(kv.first)); bow.push_back(kv.second); }
    std::vector<int> feat;                         // node id, count, indices ... in map order
    for (auto& kv : fv) { feat.push_back(static_cast<int>(kv.first)); feat.push_back(static_cast<int>(kv.second.size())); for (unsigned int i : kv.second) feat.push_back(static_cast<int>(i)); }
    const std::string pre = argv[6];
    spit(pre + ".bow", bow);
    spit(pre + ".fv", feat);
    std::vector<int> meta = {voc.getBranchingFactor(), voc.getDepthLevels(), static_cast<int>(voc.size())};
    spit(pre + ".meta", meta);
    return 0;
  }
  if (mode == "searches" && argc >= 4) {
    const Bundle b(argv[2]);
    ContextOwner ex;
    std::vector<float> dA = b.get<float>("dA"), dB = b.get<float>("dB");
    const int nA = static_cast<int>(dA.size() / 64), nB = static_cast<int>(dB.size() / 64);
    cv::Mat A(nA, 64, CV_32F, dA.data()), B(nB, 64, CV_32F, dB.data());
    const auto fvA = featvec(b.get<int>("nodeA")), fvB = featvec(b.get<int>("nodeB"));
    std::vector<int> out;
    auto put = [&](int n, const std::vector<int>& v) { out.push_back(n); out.push_back(static_cast<int>(v.size())); out.insert(out.end(), v.begin(), v.end()); };
    {  // SearchByBoW(KeyFrame*, Frame&)
      ORB_SLAM3::XFBmatcher m(ex.context(), b.scalar("ratio_kf_f"), true);
      std::vector<int> mf;
      put(m.SearchByBoW(A, fvA, b.flags("goodA"), B, fvB, mf), mf);
    }
    {  // SearchByBoW(KeyFrame*, KeyFrame*)
      ORB_SLAM3::XFBmatcher m(ex.context(), b.scalar("ratio_kf_kf"), true);
      std::vector<int> m12;
      put(m.SearchByBoW(A, fvA, b.flags("goodA"), B, fvB, b.flags("goodB"), m12), m12);
    }
    {  // SearchForTriangulation
      ORB_SLAM3::XFBmatcher m(ex.context(), 0.6f, false);
      const auto F = b.get<float>("F12"), ep = b.get<float>("ep");
      for (int coarse = 0; coarse < 2; ++coarse) {
        std::vector<std::pair<size_t, size_t> > pairs;
        const int n = m.SearchForTriangulation(A, fvA, b.flags("hasmpA"), b.flags("stereoA"), keypoints(b.get<float>("kA")), B, fvB, b.flags("hasmpB"),
                                               b.flags("stereoB"), keypoints(b.get<float>("kB")), F.data(), cv::Point2f(ep[0], ep[1]), pairs, false,
                                               coarse != 0);
        std::vector<int> m12(nA, -1);
        for (auto& pr : pairs) m12[pr.first] = static_cast<int>(pr.second);
        put(n, m12);
      }
    }
    {  // SearchByProjection(Frame&, vpMapPoints)
      ORB_SLAM3::XFBmatcher m(ex.context(), b.scalar("ratio_proj"), true);
      std::vector<float> dM = b.get<float>("dM"), dF = b.get<float>("dF");
      const int nM = static_cast<int>(dM.size() / 64), nF = static_cast<int>(dF.size() / 64);
      cv::Mat M(nM, 64, CV_32F, dM.data()), Fd(nF, 64, CV_32F, dF.data());
      const auto proj = b.get<float>("proj"), projxr = b.get<float>("projxr"), viewcos = b.get<float>("viewcos");
      const auto level = b.get<int>("level");
      const auto in_view = b.flags("in_view"), mp_obs = b.flags("mp_obs");
      std::vector<ORB_SLAM3::XFBmatcher::ProjectedPoint> pts(nM);
      for (int i = 0; i < nM; ++i) pts[i] = {in_view[i], proj[2 * i], proj[2 * i + 1], projxr[i], level[i], viewcos[i], mp_obs[i]};
      std::vector<int> assigned;
      const auto wh = b.get<float>("img_wh");
      put(m.SearchByProjection(pts, M, keypoints(b.get<float>("kF")), Fd, b.flags("occupied"), b.get<float>("uright"), 0.f, 0.f, wh[0], wh[1], 1.2f,
                               b.scalar("th_proj"), assigned), assigned);
    }
    if (b.a.count("lf_uv")) {  // SearchByProjection(CurrentFrame, LastFrame): last frame = A, current frame = B
      ORB_SLAM3::XFBmatcher m(ex.context(), 0.9f, true);
      const auto uv = b.get<float>("lf_uv"), invzc = b.get<float>("lf_invzc");
      const auto oct = b.get<int>("lf_octave");
      const auto valid = b.flags("lf_valid"), obs = b.flags("lf_obs");
      std::vector<ORB_SLAM3::XFBmatcher::LastFramePoint> pts(nA);
      for (int i = 0; i < nA; ++i) pts[i] = {valid[i], uv[2 * i], uv[2 * i + 1], invzc[i], oct[i], obs[i]};
      const auto wh = b.get<float>("img_wh");
      const auto mode3 = b.get<int>("lf_modes");   // (forward, backward) pairs to run
      for (size_t k = 0; k + 1 < mode3.size(); k += 2) {
        std::vector<int> assigned;
        put(m.SearchByProjection(pts, A, keypoints(b.get<float>("kB")), B, b.flags("occupied"), b.get<float>("uright"), 0.f, 0.f, wh[0], wh[1], 1.2f,
                                 b.scalar("lf_th"), b.scalar("lf_mbf"), mode3[k] != 0, mode3[k + 1] != 0, assigned), assigned);
      }
    }
    if (b.a.count("wq_uv")) {  // SearchByProjection(KeyFrame*, Sim3, ...) and Fuse: map points = rows of dM, keyframe = B
      ORB_SLAM3::XFBmatcher m(ex.context(), 0.6f, true);
      std::vector<float> dM = b.get<float>("dM");
      const int nM = static_cast<int>(dM.size() / 64);
      cv::Mat M(nM, 64, CV_32F, dM.data());
      const auto uv = b.get<float>("wq_uv"), ur = b.get<float>("wq_ur"), rad = b.get<float>("wq_radius");
      const auto lvl = b.get<int>("wq_level");
      const auto valid = b.flags("wq_valid");
      std::vector<ORB_SLAM3::XFBmatcher::WindowQuery> qs(nM);
      for (int i = 0; i < nM; ++i) qs[i] = {valid[i], uv[2 * i], uv[2 * i + 1], ur[i], rad[i], lvl[i]};
      const auto wh = b.get<float>("img_wh");
      std::vector<int> assigned;
      put(m.SearchByProjection(qs, M, keypoints(b.get<float>("kB")), B, b.flags("occupied"), 0.f, 0.f, wh[0], wh[1], b.scalar("wq_ratio"), assigned), assigned);
      put(m.SearchByProjectionReloc(qs, M, keypoints(b.get<float>("kB")), B, b.flags("occupied"), 0.f, 0.f, wh[0], wh[1], 64, assigned), assigned);
      std::vector<int> bi, bd;
      const std::vector<float> sig(8, b.scalar("wq_invsigma2"));
      m.FuseSearch(qs, M, keypoints(b.get<float>("kB")), b.get<float>("uright"), sig, B, 0.f, 0.f, wh[0], wh[1], true, bi, bd);
      put(0, bi); put(0, bd);
    }
    if (b.a.count("s3_uv1")) {  // SearchBySim3: keyframe 1 = A, keyframe 2 = B, map-point descriptors = the features' own rows
      ORB_SLAM3::XFBmatcher m(ex.context(), 0.6f, true);
      auto build = [&](const std::string& s, int n) {
        const auto uv = b.get<float>("s3_uv" + s), rad = b.get<float>("s3_rad" + s);
        const auto lvl = b.get<int>("s3_lvl" + s);
        const auto valid = b.flags("s3_ok" + s);
        std::vector<ORB_SLAM3::XFBmatcher::WindowQuery> q(n);
        for (int i = 0; i < n; ++i) q[i] = {valid[i], uv[2 * i], uv[2 * i + 1], 0.f, rad[i], lvl[i]};
        return q;
      };
      const auto wh = b.get<float>("img_wh");
      std::vector<int> m12;
      put(m.SearchBySim3(build("1", nA), A, keypoints(b.get<float>("kA")), A, build("2", nB), B, keypoints(b.get<float>("kB")), B, 0.f, 0.f, wh[0], wh[1],
                         m12), m12);
    }
    {  // MapPoint::ComputeDistinctiveDescriptors (batched)
      ORB_SLAM3::XFBmatcher m(ex.context(), 0.6f, true);
      std::vector<float> dS = b.get<float>("dS");
      cv::Mat S(static_cast<int>(dS.size() / 64), 64, CV_32F, dS.data());
      put(0, m.ComputeDistinctiveDescriptors(S, b.get<int>("offsets")));
    }
    spit(argv[3], out);
    return 0;
  }
  std::cerr << "usage: extract ... | init ... | bow ... | searches ..." << std::endl;
  return 1;
}
