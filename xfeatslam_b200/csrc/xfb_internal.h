// xfb_internal.h -- shared declarations of libxfeat_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "xfeat_b200.h"

namespace xfb {

// ---------------------------------------------------------------------------------------------
// Layer table.  Every BasicLayer of the reference (src/XFeat.cc:41-90) in execution order.
// ---------------------------------------------------------------------------------------------
enum LayerId {
  L_B1_0 = 0, L_B1_1, L_B1_2, L_B1_3,
  L_B2_0, L_B2_1,
  L_B3_0, L_B3_1, L_B3_2,
  L_B4_0, L_B4_1, L_B4_2,
  L_B5_0, L_B5_1, L_B5_2, L_B5_3,
  L_F_0, L_F_1,
  L_HM_0, L_HM_1,
  L_KP_0, L_KP_1, L_KP_2,
  L_NUM_BN,           // 23 BasicLayers (conv -> train-mode BN -> ReLU)
  L_F_2 = L_NUM_BN,   // block_fusion.2   1x1 + bias
  L_HM_2,             // heatmap_head.2   1x1 64->1 + bias (+ sigmoid)
  L_KP_3,             // keypoint_head.3  1x1 64->65 + bias
  L_SKIP,             // skip1.1          1x1 1->24 + bias
  L_NUM
};

// per-frame atomic counters are spaced one cache line apart: a batch's CTAs would otherwise serialise on ONE L2 line
constexpr int XFB_TICKET_STRIDE = 32;

struct LayerSpec {
  const char* ref_name;  // module path in the reference / weight blob
  int cin, cout, ks, stride;
  int lvl_in, lvl_out;   // log2 of the down-sampling factor of input / output maps
};
extern const LayerSpec kLayers[L_NUM];

// Train-mode BatchNorm statistics of one BasicLayer output, per (frame, channel).
struct BnStats {
  float* mean = nullptr;  // [B][C]
  float* rstd = nullptr;  // [B][C]   1/sqrt(biased var + 1e-5)
};

// Arguments of the generic direct-convolution kernel (conv.cu).
struct ConvArgs {
  const float* in;        // [B,Hin,Win,CIN] NHWC raw conv output of the producer (or plain values)
  const float* w;         // packed [KS*KS][CIN][COUT]
  const float* bias;      // [COUT] (OUT_BIAS modes)
  float* out;             // [B,Hout,Wout,COUT] NHWC
  const float* in_mean;   // [B][CIN] producer statistics (IN_BN*)
  const float* in_rstd;
  const float* skip_avg;  // IN_BN_SKIP: avgpool4(xn) [B,Hin,Win]
  const float* skip_w;    // [24]
  const float* skip_b;    // [24]
  double* part;           // OUT_STATS: per-block partial sums [B][tiles][COUT][2]
  unsigned int* ticket;   // [B]
  float* out_mean;        // [B][COUT]
  float* out_rstd;
  int Hin, Win, Hout, Wout;
  int full_w;             // IN_UNFOLD: width of xn
  const float* w0;        // conv_small FUSE0: block1.0's packed weights [9][1][4]
};

struct Ctx {
  int device = 0;
  int max_h = 0, max_w = 0, max_batch = 0, max_topk = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  long launches = 0;

  // weights (device): packed [KS*KS][CIN][COUT] per layer, biases where present
  float* w[L_NUM] = {};
  float* bias[L_NUM] = {};
  float* wimg[L_NUM] = {};        // tensor-core layers: per-tap hi/lo UMMA operand images (conv_tc.cu)
  unsigned char* wimg2[L_NUM] = {};   // persistent kernel (conv_tc2.cu): fp16 hi/lo operand images, resident in shared memory
  float wscale2[L_NUM] = {};      // its epilogue factor 1 / (activation scale * weight scale)
  int conv_tc_version = 2;        // 2 = conv_tc2.cu (default), 1 = conv_tc.cu (XFB_CONV_TC=1: A/B reference)
  int num_sms = 148;
  bool pdl = true;                // conv_tc2 layers launch with programmatic stream serialization (XFB_PDL=0 switches it off)
  bool b1_fuse = false;           // XFB_B1_FUSE=1: block1.0 recomputed inside block1.1 (measured: 0.212 ms vs 0.201 ms for the two kernels -- a loss, kept as an A/B option)
  struct TmapSlot { alignas(64) unsigned char blob[128]; const void* ptr; int B, H, W; };   // CUtensorMap of a layer's output + what it was encoded for
  TmapSlot tmap[L_NUM] = {};
  TmapSlot tmap_in[L_NUM] = {};   // CUtensorMap of a layer's INPUT (conv_tc2.cu raw ring: one TMA box per tile and channel phase)
  unsigned long long* t2_counters = nullptr;   // XFB_T2_DEBUG: [L_NUM][32] cycle counters of CTA 0 (conv_tc2.cu), printed at xfb_destroy
  bool force_simt = false;        // debug: run every conv on the FP32 SIMT kernels (A/B parity tests)

  // geometry of the last extract call
  int B = 0, H = 0, W = 0, in_h = 0, in_w = 0;

  // activations (device, NHWC fp32, sized for max_batch x max dims)
  uint8_t* d_gray = nullptr;      // staging for host-pointer entry points
  float* xraw = nullptr;          // [B,H,W] resized, /255
  float* xn = nullptr;            // [B,H,W] instance-normalised
  float* avg4 = nullptr;          // [B,H/4,W/4]
  float* act[L_NUM] = {};         // raw conv outputs per layer (L_F_2 = feats, L_HM_2 = H1, L_KP_3 unused)
  float* pyr = nullptr;           // [B,h,w,64] x3 + up(x4) + up(x5)
  float* k1h = nullptr;           // [B,H,W] folded keypoint heat map
  BnStats bn[L_NUM_BN];
  float* in_mean = nullptr;       // [B] instance-norm stats of the input
  float* in_rstd = nullptr;
  double* part = nullptr;         // shared partial-sum scratch
  size_t part_elems = 0;
  unsigned int* ticket = nullptr; // [B * XFB_TICKET_STRIDE]: one "CTAs done" counter per frame, each in its own 128-byte line

  // post-processing
  unsigned long long* cand = nullptr;  // [B][H*W] candidate keys
  int* cand_count = nullptr;           // [B * XFB_TICKET_STRIDE] (one cache line per frame)
  int* cand_count_last = nullptr;      // [B] copy kept for xfb_debug_candidates
  // device outputs used by the host-pointer entry points
  int32_t* o_nvalid = nullptr;
  float* o_xy = nullptr;
  float* o_score = nullptr;
  float* o_desc = nullptr;

  // pipelined submissions (xfb_submit / xfb_wait)
  struct Slot {
    uint8_t* d_gray = nullptr;
    int32_t* nvalid = nullptr; float* xy = nullptr; float* score = nullptr; float* desc = nullptr;
    int32_t* m[5] = {};
    int m_cap = 0;                 // pairs the match buffers hold
    cudaEvent_t ev_h2d = nullptr, ev_comp = nullptr, ev_d2h = nullptr, ev_ext = nullptr;
    bool pending = false;
  } slots[2];
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_match = nullptr;
  cudaEvent_t ev_match_free = nullptr;

  // last extract geometry for xfb_match_frames
  int last_topk = 0;
  const int32_t* last_nvalid = nullptr;   // device
  const float* last_desc = nullptr;       // device [B][topk][64]

  // per-kernel event timing
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;       // pairs
  std::vector<int> prof_tag;
  std::vector<cudaEvent_t> prof_pool;

  // matcher scratch
  float* m_a = nullptr; float* m_b = nullptr;
  int32_t* m_ga = nullptr; int32_t* m_gb = nullptr;
  int32_t* m_out[5] = {};
  int32_t* m_matrix = nullptr;
  int m_cap = 0;
  int32_t* m_pairs_out[5] = {};   // [n_pairs][topk] staging for xfb_match_frame_pairs
  int m_pairs_cap = 0;
  // distance-matrix operand images (match_tc.cu): [set][block][a1 16 KB | a2 16 KB] bf16, norms [set][rows_padded]
  float* tc_img[2] = {};          // generic A / B sets (xfb_match, xfb_distance_matrix)
  float* tc_nrm[2] = {};
  int tc_cap = 0;                 // rows (multiple of 128) the generic images hold
  float* tc_fnrm = nullptr;       // per-frame squared norms of the last extract (xfb_match_frame*), [frame][ms_frows]
  bool tc_fvalid = false;         // the per-frame images (ms_fimg) match the last extract result
  int32_t* tc_pairs = nullptr;    // device copy of (pairs, swapped pairs)
  int tc_pairs_cap = 0;
  float* tc_dbg = nullptr;
  float* tc_nmax[2] = {};         // [1] largest squared norm of the generic A / B sets
  float* tc_fnmax = nullptr;      // [max_batch] same per frame
  // streaming matcher (match_stream.cu): fp16 operand images, 20 KB per 128-row block
  void* ms_img[2] = {};           // generic A / B sets
  int ms_cap = 0;                 // rows (multiple of 256) the generic images hold
  void* ms_fimg = nullptr;        // (unused since the mutual matcher; kept for the grouped path's layout)
  int ms_frows = 0;
  // mutual matcher (match_mutual.cu): fp16 operand images with both tails, 22 KB per 128-row block; per-pair column scratch
  void* mm_img[2] = {};           // generic A / B sets
  int mm_cap = 0;
  void* mm_fimg = nullptr;        // per-frame images of the last extract
  int mm_frows = 0;
  unsigned int* mm_colg = nullptr; unsigned long long* mm_colk = nullptr; unsigned int* mm_done = nullptr;
  int mm_col_rows = 0;            // columns per pair the scratch holds (64 pairs)
  // vocabulary tree (bow.cu): node descriptors [n][32 B], CSR children; scratch for the host-pointer entry point
  uint8_t* v_desc = nullptr; int32_t* v_start = nullptr; int32_t* v_child = nullptr; int v_nodes = 0, v_L = 0;
  int32_t* v_out = nullptr; int v_out_cap = 0;
  // keypoint geometry (geom.cu): staging for the host-pointer entry point
  float* g_buf = nullptr; size_t g_cap = 0;      // bytes
  // pair-list distances (pairs.cu): staging for the host-pointer entry point
  int32_t* p_idx = nullptr; int32_t* p_out = nullptr; int p_cap = 0;
  unsigned long long* ms_counters = nullptr;   // debug (XFB_MS_DEBUG): device counters, see MatchTcArgs
  int ms_mode = 0;
};

// error helpers -------------------------------------------------------------------------------
void set_global_error(const std::string& s);
#define XFB_CUDA_OK(ctx, expr)                                                                      \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      (ctx)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                               \
      return XFB_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)

// Arguments of the tensor-core matcher (match_tc.cu).  A "set" is one descriptor array (one frame).
struct MatchTcArgs {
  const float* imgA; const float* imgB;   // operand images of set 0; per-pair offsets below
  const float* nrmA; const float* nrmB;   // |.|^2, [set][rows_padded]
  const float* rawA; const float* rawB;   // original fp32 rows [set][set_stride] for the exact fix-up
  int32_t pairs[128];                     // up to 64 (set of A, set of B) pairs per launch, by value
  const int32_t* nA_dev; const int32_t* nB_dev;   // per-set valid counts (device) or nullptr
  const int32_t* gA; const int32_t* gB;   // group ids (single-pair mode only) or nullptr
  int nA_host, nB_host;                   // valid counts / capacities
  int rows_padded_A, rows_padded_B;       // multiples of 128
  size_t img_stride_A, img_stride_B;      // floats per set
  size_t raw_stride_A, raw_stride_B;      // floats per set
  int init;
  int out_stride;                         // entries per pair in the outputs
  int32_t* best_idx; int32_t* best_dist; int32_t* second_dist;   // any may be null
  int32_t* matrix;                        // MATRIX mode: [nA][nB] exact distances
  float* dbg_maxerr;                      // MATRIX mode + debug: max |t - 512*float(S)|
  const float* nrm_max_B;                 // [set] largest |b|^2 of each B set (bf16 error scale)
  unsigned long long* ms_counters;        // debug: [0] queue pushes, [1] verified survivors, [2] queue overflows (or nullptr)
  int ms_mode;                            // debug timing experiments (bit flags): 1 no candidate path, 2 no epilogue arithmetic, 4 no MMAs, 8 no tcgen05.ld, 16 no bulk copies
  // mutual matcher (match_mutual.cu)
  const float* nrm_max_A;                 // [set] largest |a|^2 of each A set
  int32_t* rev_idx; int32_t* rev_dist;    // column-wise best, [pair][out_stride_cols] (any may be null)
  int out_stride_cols;
  unsigned int* col_g;                    // [pair][rows_padded_B] running maximum of the column estimates (ordered uint; zero between launches)
  unsigned long long* col_k;              // [pair][rows_padded_B] exact (distance << 32 | row) column keys (all-ones between launches)
  unsigned int* pair_done;                // [pair] CTA tickets (zero between launches)
};

// profiling tags: 0..L_NUM-1 = layers, then the stages below
enum ProfTag { P_PREP_STATS = L_NUM, P_PREP_NORM, P_PYRAMID, P_HEATMAP_OUT, P_KEYPOINT_OUT, P_NMS, P_TOPK, P_DESCRIBE, P_MATCH_TILE,
               P_DIST_PAIRS, P_DIST_MATRIX, P_MATCH_PREP, P_BOW, P_NUM };
static_assert(P_NUM <= XFB_PROF_TAGS, "profile tag table");
void prof_begin(Ctx* c, int tag);
void prof_end(Ctx* c);

// kernel launchers (each returns a cudaError_t from cudaGetLastError) -------------------------
cudaError_t launch_prep(Ctx* c, const uint8_t* d_gray, size_t frame_stride, int stride);
cudaError_t launch_conv_layer(Ctx* c, int layer);       // all BasicLayers + block_fusion.2
cudaError_t launch_pyramid(Ctx* c);
cudaError_t launch_heatmap_out(Ctx* c);                  // heatmap_head.2 + sigmoid -> H1
cudaError_t launch_keypoint_out(Ctx* c);                 // keypoint_head.3 + softmax + fold -> K1h
cudaError_t launch_post(Ctx* c, int topk, float nms_thr, int32_t* d_nvalid, float* d_xy, float* d_score, float* d_desc);
cudaError_t launch_match_prep(Ctx* c, const float* desc, size_t set_stride, int n_sets, const int32_t* n_dev, int n_host, int rows_padded,
                              float* img, size_t img_set_stride, float* nrm, float* nrm_max);
cudaError_t launch_matrix_tc(Ctx* c, const MatchTcArgs& a, int row_tiles);
cudaError_t launch_distance_pairs(Ctx* c, const float* dA, int n1, const float* dB, int n2, const int32_t* d_ia, const int32_t* d_ib, int n_pairs,
                                  int32_t* d_out);
cudaError_t launch_bow_transform(Ctx* c, const float* d_desc, size_t set_stride, int n_sets, const int32_t* n_dev, int n_host, int levelsup,
                                 int32_t* d_leaf, int32_t* d_nid, int out_stride);
void image_bounds_host(xfb_camera* cam, int w, int h);
cudaError_t launch_keypoint_geometry(Ctx* c, const float* d_xy, int n, const float* d_depth, int h, int w, int depth_stride, const xfb_camera& cam,
                                     float* d_un, float* d_depth_out, float* d_uright, int32_t* d_cell);
size_t ms_image_bytes(int rows_padded);
cudaError_t launch_ms_prep(Ctx* c, const float* desc, size_t set_stride, int n_sets, const int32_t* n_dev, int n_host, int rows_padded,
                           void* img, size_t img_set_bytes, float* nrm, float* nrm_max);
cudaError_t launch_match_stream(Ctx* c, const MatchTcArgs& a, int n_pairs, bool grouped);   // a.img_stride_* in BYTES
size_t mm_image_bytes(int rows_padded);
cudaError_t launch_mm_prep(Ctx* c, const float* desc, size_t set_stride, int n_sets, const int32_t* n_dev, int n_host, int rows_padded,
                           void* img, size_t img_set_bytes, float* nrm, float* nrm_max);
cudaError_t launch_match_mutual(Ctx* c, const MatchTcArgs& a, int n_pairs, bool mutual);     // a.img_stride_* in BYTES
size_t conv_part_elems(int H, int W);
size_t conv_tc_part_elems(int H, int W);
size_t conv_small_part_elems(int H, int W);
bool conv_small_handles(int layer);
cudaError_t launch_conv_small_layer(Ctx* c, int layer);
bool conv_tc_handles(int layer);
void conv_tc_pack_weights(int layer, const float* oihw, int cout, int cin, int ks, std::vector<float>& img);
cudaError_t launch_conv_tc_layer(Ctx* c, int layer);  // partial-sum scratch (doubles) needed per frame
size_t conv_tc2_part_floats(int H, int W);
float conv_tc2_pack_weights(int layer, const float* oihw, int cout, int cin, int ks, std::vector<unsigned char>& img);
cudaError_t launch_conv_tc2_layer(Ctx* c, int layer);

}  // namespace xfb

struct xfb_ctx : public xfb::Ctx {};
