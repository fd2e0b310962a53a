/* xfeat_b200.h -- C-ABI of libxfeat_b200.so: the B200-native (sm_100a) XFeat front-end.
 *
 * This is the drop-in boundary for xfeatSLAM's XFeat hot path.  The
 * reference has no C ABI today: its host code calls libtorch C++ directly.  Each entry point below
 * names the reference interface it replaces; the reference-side binding a maintainer would add
 * (a replacement XFextractor.cc / ORBmatcher patch) is in INTEGRATION.md and
 * xfeatslam_b200/host/.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * xfb_status; nothing throws or calls exit(); all buffers are caller-owned; one xfb_ctx per GPU
 * per host thread (a ctx is not re-entrant, like the reference's XFextractor).  There is NO CPU
 * fallback: without a CUDA device xfb_create fails with XFB_ERR_CUDA.
 */
#ifndef XFEAT_B200_H_
#define XFEAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xfb_ctx xfb_ctx;

typedef enum xfb_status {
  XFB_OK = 0,
  XFB_ERR_ARG = -1,      /* bad argument (null pointer, size out of the range given at create) */
  XFB_ERR_CUDA = -2,     /* CUDA runtime error; text in xfb_last_error */
  XFB_ERR_WEIGHTS = -3,  /* malformed / incomplete weight blob */
  XFB_ERR_EMPTY = -4,    /* empty image (the reference returns -1, src/XFextractor.cc:253-254) */
  XFB_ERR_NOMEM = -5
} xfb_status;

#define XFB_DESC_DIM 64

/* Replaces XFextractor::XFextractor (src/XFextractor.cc:75-149: weight load :133-137, device
 * select :141-144).  weights_blob/n: the flat blob written by tools/convert_weights.py from the
 * reference's weights/xfeat.pt (copied, the caller may free it after the call).  device: CUDA
 * ordinal.  max_h/max_w: largest input image; max_batch: frames per xfb_extract_batch call;
 * max_topk: largest nfeatures (<= 8192). */
int xfb_create(xfb_ctx** out, const void* weights_blob, size_t n, int device, int max_h, int max_w, int max_batch,
               int max_topk);
void xfb_destroy(xfb_ctx* ctx);

/* Last error text of this ctx (or of the last failed xfb_create when ctx == NULL). */
const char* xfb_last_error(const xfb_ctx* ctx);

/* Run all work of this ctx on an existing cudaStream_t (NULL = the ctx's own stream). */
int xfb_set_stream(xfb_ctx* ctx, void* cuda_stream);

/* Replaces the body of XFextractor::operator() up to the host packing loop
 * (src/XFextractor.cc:250-316: parseInput, preprocessTensor, XFeatModel::forward, normalize,
 * getKptsHeatmap, NMS, score, top-k, descriptor sampling, L2-normalise, valid filter).
 *   gray   : 8-bit single channel image, h rows of `stride` bytes (CV_8UC1, :257)
 *   topk   : nfeatures;  nms_thr: 0.05 in the reference (:277)
 * Outputs, sorted by score descending (ties: row-major pixel index ascending):
 *   n_valid[1]        number of keypoints with score > 0 among the top-k (<= topk)
 *   kpt_xy[topk*2]    (x, y) integer pixel coordinates in the internally resized
 *                     (multiple-of-32) frame, as float -- the reference never rescales them
 *                     (src/XFextractor.cc:304-305 multiplies int64 by trunc(ratio) = 1)
 *   score[topk], desc[topk*64]; entries >= n_valid are zero.
 * Synchronous: results are in the host buffers on return. */
int xfb_extract(xfb_ctx* ctx, const uint8_t* gray, int h, int w, int stride, int topk, float nms_thr, int32_t* n_valid,
                float* kpt_xy, float* score, float* desc);

/* `batch` frames of identical size per call; frame i starts at gray + i*frame_stride.  BatchNorm /
 * InstanceNorm statistics stay per frame (the reference is batch-1, src/XFeat.cc:19 train mode),
 * so results equal `batch` xfb_extract calls.  Output arrays are [batch] / [batch*topk*..]. */
int xfb_extract_batch(xfb_ctx* ctx, const uint8_t* gray, int batch, size_t frame_stride, int h, int w, int stride, int topk,
                      float nms_thr, int32_t* n_valid, float* kpt_xy, float* score, float* desc);

/* Same with every pointer in DEVICE memory; asynchronous on the ctx stream (no host sync). */
int xfb_extract_batch_device(xfb_ctx* ctx, const uint8_t* d_gray, int batch, size_t frame_stride, int h, int w, int stride,
                             int topk, float nms_thr, int32_t* d_n_valid, float* d_kpt_xy, float* d_score, float* d_desc);

/* Replaces ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2242-2250) for all pairs:
 * out[i*n2 + j] = int(float(||A_i - B_j||^2) * 512), A [n1,64], B [n2,64] fp32 rows
 * (cv::Mat CV_32F descriptor rows).  Bit-exact w.r.t. oracle/matcher_oracle.c for FINITE rows of any
 * norm whose distance fits an int (the tensor-core shortcut's error bound scales with |a||b| per pair;
 * large norms simply take the exact path more often).  XFeat descriptors are unit vectors. */
int xfb_distance_matrix(xfb_ctx* ctx, const float* A, int n1, const float* B, int n2, int32_t* out);
int xfb_distance_matrix_device(xfb_ctx* ctx, const float* d_A, int n1, const float* d_B, int n2, int32_t* d_out);

/* The same distance for an explicit list of pairs: out[p] = DescriptorDistance(A[idx_a[p]], B[idx_b[p]]).  This serves the
 * DescriptorDistance call sites whose candidate set is not "all pairs": the vocabulary-node gated scans of SearchByBoW
 * (src/ORBmatcher.cc:468,:489,:1019) and SearchForTriangulation (:1200), the projected-window searches (:100,:174,:699,:812,
 * :1487,:1612,:1745,:1825,:1946,:2012,:2142), Frame.cc:1079 and MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:377).
 * The caller lists the pairs in the reference's visiting order and replays the accept / reject logic over `out`
 * (xfeatslam_b200/host/XFBmatcher.cc).  Indices out of range -> XFB_ERR_ARG.  Bit-exact w.r.t. oracle/matcher_oracle.c. */
int xfb_distance_pairs(xfb_ctx* ctx, const float* A, int n1, const float* B, int n2, const int32_t* idx_a, const int32_t* idx_b, int n_pairs,
                       int32_t* out);
/* Device pointers, asynchronous on the ctx stream (out-of-range pairs produce -1). */
int xfb_distance_pairs_device(xfb_ctx* ctx, const float* d_A, int n1, const float* d_B, int n2, const int32_t* d_idx_a, const int32_t* d_idx_b,
                              int n_pairs, int32_t* d_out);

/* Brute-force nearest / second-nearest search with the reference's update rule
 * (src/ORBmatcher.cc:476-486, :884-894; lowest index wins ties), fused with the distance.
 *   group_a/group_b : nullable; when given only pairs with equal ids compete (the vocabulary-node
 *                     gating of SearchByBoW / SearchForTriangulation, src/ORBmatcher.cc:430-436)
 *   init_dist       : initial best/second value (256 in SearchByBoW :450, INT_MAX in
 *                     SearchForInitialization :860)
 * Outputs: best_idx[n1] (-1 = none), best_dist[n1], second_dist[n1]; best_idx_rev[n2] /
 * best_dist_rev[n2] = the same argmin taken column-wise (for mutual-NN checks, the spec of the
 * commented-out ORBmatcher::match, src/ORBmatcher.cc:340-406).  Any output may be NULL.
 * Rows are finite fp32 of any norm: the fp16 tensor-core filter is used while |row|^2 < 1e5 (XFeat
 * descriptors are unit vectors); a set with a larger row is matched on the exact path alone (slow,
 * still bit-exact). */
int xfb_match(xfb_ctx* ctx, const float* A, int n1, const float* B, int n2, const int32_t* group_a, const int32_t* group_b,
              int init_dist, int32_t* best_idx, int32_t* best_dist, int32_t* second_dist, int32_t* best_idx_rev,
              int32_t* best_dist_rev);
int xfb_match_device(xfb_ctx* ctx, const float* d_A, int n1, const float* d_B, int n2, const int32_t* d_group_a,
                     const int32_t* d_group_b, int init_dist, int32_t* d_best_idx, int32_t* d_best_dist, int32_t* d_second_dist,
                     int32_t* d_best_idx_rev, int32_t* d_best_dist_rev);

/* Matches the descriptors of two frames of the LAST xfb_extract_batch[_device] call without leaving
 * the device (descriptors stay resident in HBM): rows = the n_valid keypoints of frame_a, columns =
 * those of frame_b.  Output arrays are HOST buffers of `topk` entries each (entries >= n_valid are
 * -1 / init_dist); any may be NULL.  This is the fused form of "extract, then ORBmatcher on the
 * pair" used by bench.py's end-to-end leg. */
int xfb_match_frames(xfb_ctx* ctx, int frame_a, int frame_b, int init_dist, int32_t* best_idx, int32_t* best_dist,
                     int32_t* second_dist, int32_t* best_idx_rev, int32_t* best_dist_rev);
/* n_pairs matches in one call: pairs[2*p], pairs[2*p+1] = (frame_a, frame_b) of pair p (host array);
 * outputs are [n_pairs][topk] HOST arrays (one synchronisation at the end).  The _device form writes
 * to DEVICE arrays and is asynchronous on the ctx stream. */
int xfb_match_frame_pairs(xfb_ctx* ctx, const int32_t* pairs, int n_pairs, int init_dist, int32_t* best_idx, int32_t* best_dist,
                          int32_t* second_dist, int32_t* best_idx_rev, int32_t* best_dist_rev);
int xfb_match_frame_pairs_device(xfb_ctx* ctx, const int32_t* pairs, int n_pairs, int init_dist, int32_t* d_best_idx,
                                 int32_t* d_best_dist, int32_t* d_second_dist, int32_t* d_best_idx_rev, int32_t* d_best_dist_rev);

/* ---- vocabulary tree walk ---------------------------------------------------
 * Replaces the per-feature walk inside TemplatedVocabulary<FORB::TDescriptor, FORB>::transform
 * (thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1260, called from Frame::ComputeBoW src/Frame.cc:931-938 and
 * KeyFrame::ComputeBoW) with FORB::distance (FORB.cpp:81-101) -- Hamming distance over the first 32 BYTES of the
 * 64-float descriptor row, which is what the reference feeds DBoW2 for XFeat frames.
 *
 * xfb_vocab_load copies the tree: node 0 = root; the children of node i are child_index[child_start[i] .. child_start[i+1])
 * in the order of m_nodes[i].children (loadFromTextFile, TemplatedVocabulary.h:1338-1420: the order of the file);
 * node_desc[i] = the node's 32-byte ORB word; n_children == n_nodes - 1; L = m_L.
 * xfb_bow_transform: leaf[i] = the node the walk ends in (word_id / weight are m_nodes[leaf].word_id / .weight on the
 * host), nid[i] = the node passed at level L - levelsup (0 = root when L - levelsup <= 0; -1 if a leaf came first):
 * BowVector / FeatureVector are then assembled on the host exactly like transform() (:1147-1190) does
 * (xfeatslam_b200/host/XFBvocabulary.cc). */
int xfb_vocab_load(xfb_ctx* ctx, const uint8_t* node_desc, const int32_t* child_start, const int32_t* child_index, int n_nodes, int n_children,
                   int L);
int xfb_bow_transform(xfb_ctx* ctx, const float* desc, int n, int levelsup, int32_t* leaf, int32_t* nid);
int xfb_bow_transform_device(xfb_ctx* ctx, const float* d_desc, int n, int levelsup, int32_t* d_leaf, int32_t* d_nid);
/* Every frame of the last xfb_extract_batch[_device] call in one launch (descriptors stay in HBM); outputs are DEVICE arrays
 * [batch][topk], entries >= n_valid are -1; asynchronous on the ctx stream. */
int xfb_bow_transform_frames_device(xfb_ctx* ctx, int levelsup, int32_t* d_leaf, int32_t* d_nid);

/* ---- per-keypoint frame geometry -------------------------------------------
 * What the reference's Frame constructor computes for every keypoint right after extraction (src/Frame.cc:289-345):
 * Frame::UndistortKeyPoints (:940-973, cv::undistortPoints with K, mDistCoef, P = mK), Frame::ComputeStereoFromRGBD
 * (:1177-1198) and the cell of Frame::AssignFeaturesToGrid / PosInGrid (:569-600, :918-928) -- one launch per frame. */
typedef struct xfb_camera {
  float fx, fy, cx, cy;               /* Pinhole::toK() == mK */
  float k1, k2, p1, p2, k3;           /* mDistCoef (k3 = 0 for 4 coefficients); k1 == 0: keypoints are copied, as :942-946 */
  float bf;                           /* mbf (stereo baseline times fx) */
  float min_x, min_y, max_x, max_y;   /* mnMinX, mnMinY, mnMaxX, mnMaxY: filled by xfb_image_bounds */
} xfb_camera;
/* Frame::ComputeImageBounds (src/Frame.cc:975-1003): undistorted image corners -> cam->min_x .. max_y.  Host only, no ctx. */
int xfb_image_bounds(xfb_camera* cam, int w, int h);
/* xy [n,2]: keypoint positions as XFextractor packs them (mvKeys[i].pt, phantom (0,0) rows included); depth: the CV_32F depth
 * image in metres (after mDepthMapFactor), `depth_stride` floats per row, NULL for monocular frames.  Outputs (any may be
 * NULL): un_xy [n,2] = mvKeysUn[i].pt; kp_depth [n] = mvDepth; uright [n] = mvuRight (both -1 without a positive depth);
 * cell [n] = posX * 48 + posY of mGrid[posX][posY], -1 when PosInGrid fails.  Host pointers, synchronous. */
int xfb_keypoint_geometry(xfb_ctx* ctx, const float* xy, int n, const float* depth, int h, int w, int depth_stride, const xfb_camera* cam,
                          float* un_xy, float* kp_depth, float* uright, int32_t* cell);
/* Device pointers, asynchronous on the ctx stream. */
int xfb_keypoint_geometry_device(xfb_ctx* ctx, const float* d_xy, int n, const float* d_depth, int h, int w, int depth_stride,
                                 const xfb_camera* cam, float* d_un_xy, float* d_kp_depth, float* d_uright, int32_t* d_cell);

/* ---- pipelined form for frame streams ------------------------------------------------------
 * xfb_submit enqueues one batch: host->device copy of the frames, extraction, (optionally) the frame-pair
 * matches of xfb_match_frame_pairs, and the device->host copies of every requested output -- on three CUDA
 * streams (copy-in, compute, copy-out) chained by events, and returns immediately.  Two slots (0, 1) can
 * be in flight, so the copies of batch i+1 / i-1 overlap the kernels of batch i.  xfb_wait blocks until
 * the slot's outputs are in the caller's buffers.  Host buffers should be pinned (cudaHostAlloc /
 * cudaHostRegister) for the copies to be truly asynchronous; all buffers stay caller-owned and must
 * remain valid until xfb_wait(slot) returns.  Match outputs are [n_pairs][topk]; any may be NULL. */
int xfb_submit(xfb_ctx* ctx, int slot, const uint8_t* gray, int batch, size_t frame_stride, int h, int w, int stride, int topk,
               float nms_thr, int32_t* n_valid, float* kpt_xy, float* score, float* desc, const int32_t* pairs, int n_pairs,
               int init_dist, int32_t* best_idx, int32_t* best_dist, int32_t* second_dist, int32_t* best_idx_rev,
               int32_t* best_dist_rev);
int xfb_wait(xfb_ctx* ctx, int slot);

/* ---- introspection (used by the parity tests and bench.py) ------------------------------- */

/* Copies an intermediate of the last extract call to the host as fp32.  `name` is a reference
 * module path ("block1.0", ..., "block5.3", "block_fusion.0", "heatmap_head.1", ... = the RAW conv
 * output of that BasicLayer, NHWC; "xn", "pyramid_sum", "feats", "H1", "K1h").  dims[4] receives
 * {H, W, C, 0}.  Returns the element count, or a negative xfb_status. */
long xfb_debug_read(xfb_ctx* ctx, const char* name, int frame, float* host_out, size_t capacity, int32_t* dims);
/* Per-(frame, channel) batch statistics (mean, 1/sqrt(var+eps)) of a BasicLayer; out [2*C]. */
long xfb_debug_read_stats(xfb_ctx* ctx, const char* name, int frame, float* host_out, size_t capacity);
/* Overwrites a dense map ("feats" [h,w,64] NHWC, "H1" [h,w], "K1h" [H,W]) of frame 0 and re-runs
 * only the discrete post-processing (NMS, score, top-k, descriptors) on it, for bit-exact tests
 * on oracle-provided inputs.  Inputs are host fp32; outputs as in xfb_extract. */
int xfb_debug_post(xfb_ctx* ctx, int H, int W, const float* feats, const float* H1, const float* K1h, int topk, float nms_thr,
                   int32_t* n_valid, float* kpt_xy, float* score, float* desc);
/* Largest |t - 512*float(||a-b||^2)| over all pairs, where t is the tensor-core estimate (bf16 two-piece split,
 * a1.b1 + a1.b2 + a2.b1) that xfb_distance_matrix takes floor() of: must stay below the kernel's per-pair bound
 * eps = 0.03 |a||b| + 0.01 (|a|^2 + |b|^2) + 0.005 (csrc/match_tc.cu match_eps) for the shortcut to be sound. */
int xfb_debug_match_error(xfb_ctx* ctx, const float* A, int n1, const float* B, int n2, float* max_err);
/* Number of NMS candidates (score > 0) of `frame` in the last extract call. */
int xfb_debug_candidates(xfb_ctx* ctx, int frame);
/* Debug A/B switch: enable != 0 runs every convolution on the FP32 SIMT kernels instead of the tcgen05
 * implicit-GEMM kernels (used by the parity tests to compare the two; both are sm_100a CUDA). */
int xfb_debug_force_simt(xfb_ctx* ctx, int enable);
/* Total number of kernels this ctx has launched so far. */
long xfb_launch_count(const xfb_ctx* ctx);
/* Per-kernel CUDA-event timing on the ctx stream.  enable != 0 starts recording an event pair around
 * every kernel launch; xfb_profile_read synchronises, adds the elapsed times into ms[tag] / count[tag]
 * (arrays of XFB_PROF_TAGS entries, caller-zeroed) and clears the recorded events.
 * xfb_profile_tag_name(tag) names a tag (a reference module path or a stage name). */
#define XFB_PROF_TAGS 40
int xfb_profile_enable(xfb_ctx* ctx, int enable);
int xfb_profile_read(xfb_ctx* ctx, float* ms, int32_t* count);
const char* xfb_profile_tag_name(int tag);

#ifdef __cplusplus
}
#endif
#endif /* XFEAT_B200_H_ */
