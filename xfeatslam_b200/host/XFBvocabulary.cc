// XFBvocabulary.cc -- see XFBvocabulary.h.
#include "XFBvocabulary.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace ORB_SLAM3 {

XFBvocabulary::XFBvocabulary(xfb_ctx* ctx, int k, int L, ScoringType scoring, WeightingType weighting, const std::vector<unsigned char>& node_desc,
                             const std::vector<int32_t>& child_start, const std::vector<int32_t>& child_index, const std::vector<double>& weight,
                             const std::vector<int32_t>& word_id)
    : ctx_(ctx), m_k(k), m_L(L), m_scoring(scoring), m_weighting(weighting), weight_(weight), word_id_(word_id), m_words(0) {
  if (!ctx) throw std::invalid_argument("XFBvocabulary: null xfb_ctx (no CPU fallback)");
  const int n_nodes = static_cast<int>(weight.size());
  if (node_desc.size() != static_cast<size_t>(n_nodes) * 32 || child_start.size() != static_cast<size_t>(n_nodes) + 1 ||
      word_id.size() != static_cast<size_t>(n_nodes))
    throw std::invalid_argument("XFBvocabulary: inconsistent array sizes");
  for (int32_t w : word_id_) if (w >= 0) m_words++;
  if (xfb_vocab_load(ctx_, node_desc.data(), child_start.data(), child_index.data(), n_nodes, static_cast<int>(child_index.size()), L) != XFB_OK)
    throw std::runtime_error(std::string("xfb_vocab_load: ") + xfb_last_error(ctx_));
}

XFBvocabulary XFBvocabulary::loadFromTextFile(xfb_ctx* ctx, const std::string& filename) {
  std::ifstream f(filename.c_str());
  if (!f) throw std::runtime_error("XFBvocabulary: cannot open " + filename);
  std::string s;
  std::getline(f, s);
  std::stringstream ss(s);
  int k, L, n1, n2;
  ss >> k >> L >> n1 >> n2;
  if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3)   // TemplatedVocabulary.h:1360-1364
    throw std::runtime_error("Vocabulary loading failure: This is not a correct text file!");
  std::vector<int> parent(1, 0);
  std::vector<unsigned char> desc(32, 0);          // node 0 = root, no word
  std::vector<double> weight(1, 0.0);
  std::vector<int32_t> word_id(1, -1);
  int words = 0;
  while (std::getline(f, s)) {
    if (s.empty()) continue;
    std::stringstream sn(s);
    int pid, leaf;
    sn >> pid >> leaf;
    parent.push_back(pid);
    for (int i = 0; i < 32; ++i) { int b; sn >> b; desc.push_back(static_cast<unsigned char>(b)); }   // FORB::fromString
    double w;
    sn >> w;
    weight.push_back(w);
    word_id.push_back(leaf > 0 ? words++ : -1);
  }
  const int n = static_cast<int>(parent.size());
  // m_nodes[pid].children.push_back(nid) in file order -> CSR
  std::vector<int32_t> child_start(static_cast<size_t>(n) + 1, 0), child_index(static_cast<size_t>(n) - 1);
  for (int i = 1; i < n; ++i) child_start[parent[i] + 1]++;
  for (int i = 0; i < n; ++i) child_start[i + 1] += child_start[i];
  std::vector<int32_t> fill(child_start.begin(), child_start.end() - 1);
  for (int i = 1; i < n; ++i) child_index[fill[parent[i]]++] = i;
  return XFBvocabulary(ctx, k, L, static_cast<ScoringType>(n1), static_cast<WeightingType>(n2), desc, child_start, child_index, weight, word_id);
}

void XFBvocabulary::transform(const cv::Mat& descriptors, BowVector& v, FeatureVector& fv, int levelsup) const {
  v.clear();
  fv.clear();
  const int N = descriptors.rows;
  if (N == 0) return;
  if (descriptors.type() != CV_32F || descriptors.cols != XFB_DESC_DIM) throw std::invalid_argument("XFBvocabulary: descriptors must be CV_32F N x 64");
  std::vector<float> rows(static_cast<size_t>(N) * XFB_DESC_DIM);
  for (int r = 0; r < N; ++r) std::memcpy(rows.data() + static_cast<size_t>(r) * XFB_DESC_DIM, descriptors.ptr<float>(r), XFB_DESC_DIM * sizeof(float));
  std::vector<int32_t> leaf(N), nid(N);
  if (xfb_bow_transform(ctx_, rows.data(), N, levelsup, leaf.data(), nid.data()) != XFB_OK)
    throw std::runtime_error(std::string("xfb_bow_transform: ") + xfb_last_error(ctx_));
  // mustNormalize (thirdparty/DBoW2/DBoW2/ScoringObject.h:73-90): every scoring object but the dot product normalises,
  // L2Scoring with the L2 norm, the others with L1
  const bool must = (m_scoring != DOT_PRODUCT);
  if (m_weighting == TF || m_weighting == TF_IDF) {           // TemplatedVocabulary.h:1147-1173
    for (int i = 0; i < N; ++i) {
      const double w = weight_[leaf[i]];
      if (w > 0) {                                            // not stopped
        v[static_cast<unsigned int>(word_id_[leaf[i]])] += w; // BowVector::addWeight
        fv[static_cast<unsigned int>(nid[i])].push_back(static_cast<unsigned int>(i));
      }
    }
    if (!v.empty() && !must) {
      const double nd = static_cast<double>(v.size());
      for (BowVector::iterator it = v.begin(); it != v.end(); ++it) it->second /= nd;
    }
  } else {                                                    // IDF || BINARY (:1175-1191)
    for (int i = 0; i < N; ++i) {
      const double w = weight_[leaf[i]];
      if (w > 0) {
        v.insert(BowVector::value_type(static_cast<unsigned int>(word_id_[leaf[i]]), w));   // addIfNotExist
        fv[static_cast<unsigned int>(nid[i])].push_back(static_cast<unsigned int>(i));
      }
    }
  }
  if (must) {                                                 // BowVector::normalize, BowVector.cpp:62-84
    double norm = 0.0;
    if (m_scoring != L2_NORM) { for (BowVector::iterator it = v.begin(); it != v.end(); ++it) norm += std::fabs(it->second); }
    else { for (BowVector::iterator it = v.begin(); it != v.end(); ++it) norm += it->second * it->second; norm = std::sqrt(norm); }
    if (norm > 0.0) for (BowVector::iterator it = v.begin(); it != v.end(); ++it) it->second /= norm;
  }
}

}  // namespace ORB_SLAM3
