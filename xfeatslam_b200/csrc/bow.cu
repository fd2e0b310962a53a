// bow.cu -- the per-feature vocabulary-tree walk of DBoW2 on the device.
//
// Replaces the inner call of TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector&, FeatureVector&, levelsup)
// (thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1194), i.e. transform(feature, word_id, weight, &nid, levelsup) (:1218-1260),
// with FORB::distance (thirdparty/DBoW2/DBoW2/FORB.cpp:81-101): the Hamming distance over 8 int32 words.  For an XFeat frame the
// "feature" handed to DBoW2 is a 1 x 64 CV_32F row (Frame::ComputeBoW, src/Frame.cc:931-938, Converter::toDescriptorVector), and
// FORB::distance reads its first 32 BYTES -- the bit patterns of the first 8 floats -- against the 32-byte ORB word of each node.
// That is what the reference computes, and so does this kernel.
//
// Integer / byte work: one thread per feature walks root -> leaf (L levels, k children each: 8 XOR + POPC per child, the first
// child with the smallest distance wins, `d < best_d`), recording the node reached at level L - levelsup.  The node table
// (1.1 M nodes x 32 B for ORBvoc) is gathered through L2; the walk is latency bound (L dependent gathers), so the launch covers
// every frame of a batch at once.
#include <cuda_runtime.h>

#include "xfb_internal.h"

namespace xfb {

__global__ void __launch_bounds__(128) bow_transform_kernel(const float* __restrict__ desc, size_t set_stride, const int32_t* __restrict__ n_dev,
                                                            int n_host, const uint4* __restrict__ node_desc, const int32_t* __restrict__ child_start,
                                                            const int32_t* __restrict__ child_index, int nid_level, int32_t* __restrict__ leaf_out,
                                                            int32_t* __restrict__ nid_out, int out_stride) {
  const int set = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_dev ? min(n_host, n_dev[set]) : n_host;
  if (i >= out_stride) return;
  const size_t o = (size_t)set * out_stride + i;
  if (i >= n) { leaf_out[o] = -1; nid_out[o] = -1; return; }
  const uint4* f4 = reinterpret_cast<const uint4*>(desc + (size_t)set * set_stride + (size_t)i * 64);
  const uint4 fa = f4[0], fb = f4[1];                       // the first 32 bytes of the descriptor row
  int node = 0, level = 0;
  int nid = (nid_level <= 0) ? 0 : -1;                      // :1227-1228 (root); -1 = the walk ended above nid_level (unbalanced tree)
  int c0 = child_start[0], c1 = child_start[1];
  while (c1 > c0) {                                         // do { ... } while (!isLeaf())
    ++level;
    int best = -1, best_d = 0x7fffffff;
    for (int c = c0; c < c1; ++c) {
      const int id = child_index[c];
      const uint4 da = node_desc[2 * (size_t)id], db = node_desc[2 * (size_t)id + 1];
      const int d = __popc(fa.x ^ da.x) + __popc(fa.y ^ da.y) + __popc(fa.z ^ da.z) + __popc(fa.w ^ da.w) + __popc(fb.x ^ db.x) +
                    __popc(fb.y ^ db.y) + __popc(fb.z ^ db.z) + __popc(fb.w ^ db.w);
      if (d < best_d) { best_d = d; best = id; }            // nodes[0] first, then `if (d < best_d)`: the first minimum wins
    }
    node = best;
    if (level == nid_level) nid = node;
    c0 = child_start[node]; c1 = child_start[node + 1];
  }
  leaf_out[o] = node;
  nid_out[o] = nid;
}

cudaError_t launch_bow_transform(Ctx* c, const float* d_desc, size_t set_stride, int n_sets, const int32_t* n_dev, int n_host, int levelsup,
                                 int32_t* d_leaf, int32_t* d_nid, int out_stride) {
  if (out_stride <= 0 || n_sets <= 0) return cudaSuccess;
  dim3 grid((out_stride + 127) / 128, n_sets);
  prof_begin(c, P_BOW);
  bow_transform_kernel<<<grid, 128, 0, c->stream>>>(d_desc, set_stride, n_dev, n_host, reinterpret_cast<const uint4*>(c->v_desc), c->v_start, c->v_child,
                                                    c->v_L - levelsup, d_leaf, d_nid, out_stride);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

}  // namespace xfb
