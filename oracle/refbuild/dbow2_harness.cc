// TEST INFRASTRUCTURE ONLY -- drives the reference's OWN vendored DBoW2 (thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h + FORB.cpp,
// compiled unchanged, see Makefile target dbow2) the way Frame::ComputeBoW does (src/Frame.cc:931-938):
//   vector<cv::Mat> vCurrentDesc = Converter::toDescriptorVector(mDescriptors);   // rows of the CV_32F descriptor matrix
//   mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4);
// usage: ref_dbow2 <vocabulary.txt> <desc.f32> <n> <levelsup> <out prefix>
//   <out>.leaf  int32 [n][3]: word id, node id at level L - levelsup (single-feature transform, TemplatedVocabulary.h:1218-1260), weight > 0
//   <out>.bow   float64 pairs (word id, value)      -- BowVector after transform()
//   <out>.fv    int32 stream: node id, count, feature indices ...   -- FeatureVector
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "FORB.h"
#include "TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabularyBase;   // include/ORBVocabulary.h:29-30
struct ORBVocabulary : ORBVocabularyBase { using ORBVocabularyBase::transform; };   // the per-feature walk is protected: expose it, unchanged

int main(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage: ref_dbow2 voc.txt desc.f32 n levelsup out\n"); return 2; }
  ORBVocabulary voc;
  if (!voc.loadFromTextFile(argv[1])) { std::fprintf(stderr, "cannot load %s\n", argv[1]); return 1; }
  const int n = std::atoi(argv[3]), levelsup = std::atoi(argv[4]);
  std::vector<float> desc((size_t)n * 64);
  { std::ifstream f(argv[2], std::ios::binary); f.read(reinterpret_cast<char*>(desc.data()), desc.size() * 4); if (!f) return 1; }
  std::vector<cv::Mat> rows;
  rows.reserve(n);
  for (int j = 0; j < n; ++j) rows.push_back(cv::Mat(1, 64, CV_32F, desc.data() + (size_t)j * 64));   // Descriptors.row(j)
  const std::string out = argv[5];
  {
    std::vector<int> leaf((size_t)n * 3);
    for (int j = 0; j < n; ++j) {
      DBoW2::WordId id; DBoW2::WordValue w; DBoW2::NodeId nid;
      voc.transform(rows[j], id, w, &nid, levelsup);
      leaf[3 * j] = (int)id; leaf[3 * j + 1] = (int)nid; leaf[3 * j + 2] = w > 0 ? 1 : 0;
    }
    std::ofstream f(out + ".leaf", std::ios::binary); f.write(reinterpret_cast<const char*>(leaf.data()), leaf.size() * 4);
  }
  DBoW2::BowVector bow; DBoW2::FeatureVector fv;
  voc.transform(rows, bow, fv, levelsup);
  {
    std::ofstream f(out + ".bow", std::ios::binary);
    for (auto& kv : bow) { double p[2] = {(double)kv.first, kv.second}; f.write(reinterpret_cast<const char*>(p), 16); }
  }
  {
    std::ofstream f(out + ".fv", std::ios::binary);
    for (auto& kv : fv) {
      std::vector<int> rec; rec.push_back((int)kv.first); rec.push_back((int)kv.second.size());
      for (unsigned int i : kv.second) rec.push_back((int)i);
      f.write(reinterpret_cast<const char*>(rec.data()), rec.size() * 4);
    }
  }
  std::printf("{\"k\": %d, \"L\": %d, \"words\": %u}\n", voc.getBranchingFactor(), voc.getDepthLevels(), voc.size());
  return 0;
}
