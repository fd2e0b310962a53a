// XFBvocabulary.h -- host-side companion of DBoW2's ORBVocabulary for the XFeat path.
//
// The reference computes Frame::mBowVec / mFeatVec with
//     mpORBvocabulary->transform(Converter::toDescriptorVector(mDescriptors), mBowVec, mFeatVec, 4);   (src/Frame.cc:931-938)
// i.e. TemplatedVocabulary<FORB::TDescriptor, FORB>::transform (thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1194): one tree
// walk per feature (:1218-1260) + map bookkeeping.  Here the tree walks of ALL features run in one GPU launch
// (xfb_bow_transform, csrc/bow.cu) and the bookkeeping -- weights, stop words, BowVector accumulation / normalisation,
// FeatureVector -- is replayed on the host exactly as transform() does it.
//
// Scope: the vocabulary ORB-SLAM3 ships (ORBvoc.txt: TF_IDF weighting, L1_NORM scoring); the text loader follows
// loadFromTextFile (:1338-1420).  BowVector / FeatureVector are the std::maps DBoW2 derives from.
#ifndef XFBVOCABULARY_H
#define XFBVOCABULARY_H

#include <map>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp>

#include "xfeat_b200.h"

namespace ORB_SLAM3 {

class XFBvocabulary {
 public:
  typedef std::map<unsigned int, double> BowVector;                            // DBoW2::BowVector  (WordId -> WordValue)
  typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;    // DBoW2::FeatureVector (NodeId -> feature indices)
  enum WeightingType { TF_IDF, TF, IDF, BINARY };                              // DBoW2::WeightingType
  enum ScoringType { L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT };

  // Uploads the tree to the device of `ctx` (xfb_vocab_load).  Arrays as in include/xfeat_b200.h; weight / word_id per node.
  XFBvocabulary(xfb_ctx* ctx, int k, int L, ScoringType scoring, WeightingType weighting, const std::vector<unsigned char>& node_desc,
                const std::vector<int32_t>& child_start, const std::vector<int32_t>& child_index, const std::vector<double>& weight,
                const std::vector<int32_t>& word_id);
  // TemplatedVocabulary::loadFromTextFile
  static XFBvocabulary loadFromTextFile(xfb_ctx* ctx, const std::string& filename);

  // TemplatedVocabulary::transform(features, v, fv, levelsup) for the rows of a CV_32F N x 64 descriptor matrix
  void transform(const cv::Mat& descriptors, BowVector& v, FeatureVector& fv, int levelsup) const;

  int getBranchingFactor() const { return m_k; }
  int getDepthLevels() const { return m_L; }
  unsigned int size() const { return m_words; }

 private:
  xfb_ctx* ctx_;
  int m_k, m_L;
  ScoringType m_scoring;
  WeightingType m_weighting;
  std::vector<double> weight_;
  std::vector<int32_t> word_id_;
  unsigned int m_words;
};

}  // namespace ORB_SLAM3
#endif
