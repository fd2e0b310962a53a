// api.cu -- context management, the extract pipeline and the C-ABI of libxfeat_b200.so.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "xfb_internal.h"

namespace xfb {

static std::mutex g_err_mu;
static std::string g_err;
void set_global_error(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = s;
}

// ---- per-kernel event timing ---------------------------------------------------------------------------
static cudaEvent_t prof_get_event(Ctx* c) {
  cudaEvent_t e = nullptr;
  if (!c->prof_pool.empty()) { e = c->prof_pool.back(); c->prof_pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}
void prof_begin(Ctx* c, int tag) {
  if (!c->prof) return;
  cudaEvent_t e = prof_get_event(c);
  cudaEventRecord(e, c->stream);
  c->prof_ev.push_back(e);
  c->prof_tag.push_back(tag);
}
void prof_end(Ctx* c) {
  if (!c->prof) return;
  cudaEvent_t e = prof_get_event(c);
  cudaEventRecord(e, c->stream);
  c->prof_ev.push_back(e);
}
static const char* kStageNames[] = {"prep_stats", "prep_norm", "pyramid_sum", "heatmap_out", "keypoint_out_softmax_fold", "nms_score",
                                    "topk_select_sort", "describe", "match_tile", "distance_pairs", "distance_matrix", "match_prep", "bow_transform"};

// ---- weight blob (tools/convert_weights.py) --------------------------------------------------------
#pragma pack(push, 1)
struct BlobEntry {
  char name[48];
  uint32_t ndim;
  uint32_t dims[4];
  uint64_t offset;
  uint64_t nbytes;
};
#pragma pack(pop)
static_assert(sizeof(BlobEntry) == 48 + 4 + 16 + 8 + 8, "blob table entry layout");

static const float* blob_find(const uint8_t* blob, size_t n, const std::string& name, size_t expect_elems) {
  if (n < 16 || std::memcmp(blob, "XFBW", 4) != 0) return nullptr;
  uint32_t version, count;
  std::memcpy(&version, blob + 4, 4);
  std::memcpy(&count, blob + 8, 4);
  if (version != 1 || 16 + (size_t)count * sizeof(BlobEntry) > n) return nullptr;
  for (uint32_t i = 0; i < count; ++i) {
    BlobEntry e;
    std::memcpy(&e, blob + 16 + (size_t)i * sizeof(BlobEntry), sizeof(e));
    char nm[49];
    std::memcpy(nm, e.name, 48);
    nm[48] = 0;
    if (name == nm) {
      if (e.offset > n || e.nbytes > n - e.offset || e.nbytes != expect_elems * 4 || (e.offset & 3)) return nullptr;
      return reinterpret_cast<const float*>(blob + e.offset);
    }
  }
  return nullptr;
}

static int dev_alloc(Ctx* c, void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    c->err = std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
    return XFB_ERR_NOMEM;
  }
  return XFB_OK;
}
#define XFB_ALLOC(ctx, ptr, bytes)                                         \
  do {                                                                     \
    int _r = dev_alloc((ctx), reinterpret_cast<void**>(&(ptr)), (bytes)); \
    if (_r != XFB_OK) return _r;                                           \
  } while (0)

static int load_weights(Ctx* c, const uint8_t* blob, size_t n) {
  for (int L = 0; L < L_NUM; ++L) {
    const LayerSpec& sp = kLayers[L];
    const bool is_basic = L < L_NUM_BN;
    const std::string wname = std::string(sp.ref_name) + (is_basic ? ".layer.0.weight" : ".weight");
    const size_t kk = (size_t)sp.ks * sp.ks, elems = (size_t)sp.cout * sp.cin * kk;
    const float* src = blob_find(blob, n, wname, elems);
    if (!src) { c->err = "weight blob: missing or malformed tensor " + wname; return XFB_ERR_WEIGHTS; }
    // OIHW -> [ky*KS+kx][ci][co]
    std::vector<float> packed(elems);
    for (int co = 0; co < sp.cout; ++co)
      for (int ci = 0; ci < sp.cin; ++ci)
        for (size_t k = 0; k < kk; ++k) packed[(k * sp.cin + ci) * sp.cout + co] = src[((size_t)co * sp.cin + ci) * kk + k];
    XFB_ALLOC(c, c->w[L], elems * 4);
    XFB_CUDA_OK(c, cudaMemcpy(c->w[L], packed.data(), elems * 4, cudaMemcpyHostToDevice));
    if (conv_tc_handles(L)) {
      std::vector<float> img;
      conv_tc_pack_weights(L, src, sp.cout, sp.cin, sp.ks, img);
      XFB_ALLOC(c, c->wimg[L], img.size() * 4);
      XFB_CUDA_OK(c, cudaMemcpy(c->wimg[L], img.data(), img.size() * 4, cudaMemcpyHostToDevice));
      std::vector<unsigned char> img2;
      c->wscale2[L] = conv_tc2_pack_weights(L, src, sp.cout, sp.cin, sp.ks, img2);
      XFB_ALLOC(c, c->wimg2[L], img2.size());
      XFB_CUDA_OK(c, cudaMemcpy(c->wimg2[L], img2.data(), img2.size(), cudaMemcpyHostToDevice));
    }
    if (!is_basic) {
      const std::string bname = std::string(sp.ref_name) + ".bias";
      const float* bsrc = blob_find(blob, n, bname, sp.cout);
      if (!bsrc) { c->err = "weight blob: missing or malformed tensor " + bname; return XFB_ERR_WEIGHTS; }
      XFB_ALLOC(c, c->bias[L], (size_t)sp.cout * 4);
      XFB_CUDA_OK(c, cudaMemcpy(c->bias[L], bsrc, (size_t)sp.cout * 4, cudaMemcpyHostToDevice));
    }
  }
  return XFB_OK;
}

static int alloc_buffers(Ctx* c) {
  const size_t B = c->max_batch;
  const size_t H = (size_t)(c->max_h / 32) * 32, W = (size_t)(c->max_w / 32) * 32;
  const size_t HW = H * W;
  XFB_ALLOC(c, c->d_gray, B * (size_t)c->max_h * c->max_w);
  XFB_ALLOC(c, c->xraw, B * HW * 4);
  XFB_ALLOC(c, c->xn, B * HW * 4);
  XFB_ALLOC(c, c->avg4, B * (HW / 16) * 4);
  for (int L = 0; L < L_NUM; ++L) {
    if (L == L_KP_3 || L == L_SKIP) continue;
    const LayerSpec& sp = kLayers[L];
    const size_t px = (H >> sp.lvl_out) * (W >> sp.lvl_out);
    XFB_ALLOC(c, c->act[L], B * px * sp.cout * 4);
    if (L < L_NUM_BN) {
      XFB_ALLOC(c, c->bn[L].mean, B * sp.cout * 4);
      XFB_ALLOC(c, c->bn[L].rstd, B * sp.cout * 4);
    }
  }
  XFB_ALLOC(c, c->pyr, B * (HW / 64) * 64 * 4);
  XFB_ALLOC(c, c->k1h, B * HW * 4);
  XFB_ALLOC(c, c->in_mean, B * 4);
  XFB_ALLOC(c, c->in_rstd, B * 4);
  size_t pe = conv_part_elems((int)H, (int)W);
  if (conv_tc_part_elems((int)H, (int)W) > pe) pe = conv_tc_part_elems((int)H, (int)W);
  if (conv_small_part_elems((int)H, (int)W) > pe) pe = conv_small_part_elems((int)H, (int)W);
  if ((conv_tc2_part_floats((int)H, (int)W) + 1) / 2 > pe) pe = (conv_tc2_part_floats((int)H, (int)W) + 1) / 2;
  const size_t prep_pe = ((HW + 2047) / 2048) * 2;
  if (prep_pe > pe) pe = prep_pe;
  c->part_elems = pe;
  XFB_ALLOC(c, c->part, B * pe * sizeof(double));
  XFB_ALLOC(c, c->ticket, B * XFB_TICKET_STRIDE * 4);
  XFB_CUDA_OK(c, cudaMemset(c->ticket, 0, B * XFB_TICKET_STRIDE * 4));
  XFB_ALLOC(c, c->cand, B * HW * 8);
  XFB_ALLOC(c, c->cand_count, B * XFB_TICKET_STRIDE * 4);   // one 128-byte line per frame: the CTAs of a batch would otherwise serialise their atomics on ONE L2 line
  XFB_ALLOC(c, c->cand_count_last, B * 4);
  XFB_CUDA_OK(c, cudaMemset(c->cand_count, 0, B * XFB_TICKET_STRIDE * 4));
  XFB_CUDA_OK(c, cudaMemset(c->cand_count_last, 0, B * 4));
  const size_t K = c->max_topk;
  XFB_ALLOC(c, c->o_nvalid, B * 4);
  XFB_ALLOC(c, c->o_xy, B * K * 2 * 4);
  XFB_ALLOC(c, c->o_score, B * K * 4);
  XFB_ALLOC(c, c->o_desc, B * K * 64 * 4);
  return XFB_OK;
}

static int ensure_match_scratch(Ctx* c, int n1, int n2, bool want_matrix) {
  const int n = n1 > n2 ? n1 : n2;
  if (n > c->m_cap) {
    float** fp[2] = {&c->m_a, &c->m_b};
    for (auto p : fp) { if (*p) cudaFree(*p); *p = nullptr; }
    int32_t** ip[] = {&c->m_ga, &c->m_gb, &c->m_out[0], &c->m_out[1], &c->m_out[2], &c->m_out[3], &c->m_out[4], &c->m_matrix};
    for (auto p : ip) { if (*p) cudaFree(*p); *p = nullptr; }
    const size_t cap = (size_t)((n + 63) / 64) * 64;
    XFB_ALLOC(c, c->m_a, cap * 64 * 4);
    XFB_ALLOC(c, c->m_b, cap * 64 * 4);
    XFB_ALLOC(c, c->m_ga, cap * 4);
    XFB_ALLOC(c, c->m_gb, cap * 4);
    for (int i = 0; i < 5; ++i) XFB_ALLOC(c, c->m_out[i], cap * 4);
    c->m_cap = (int)cap;
  }
  if (want_matrix && !c->m_matrix) XFB_ALLOC(c, c->m_matrix, (size_t)c->m_cap * c->m_cap * 4);
  return XFB_OK;
}

// ---- tensor-core matcher plumbing (match_tc.cu) -------------------------------------------------------------
static inline int pad128(int n) { return n <= 0 ? 128 : (n + 127) / 128 * 128; }

static int tc_ensure_generic(Ctx* c, int n1, int n2) {
  const int need = pad128(n1 > n2 ? n1 : n2);
  if (need > c->tc_cap) {
    for (int i = 0; i < 2; ++i) { if (c->tc_img[i]) cudaFree(c->tc_img[i]); if (c->tc_nrm[i]) cudaFree(c->tc_nrm[i]); c->tc_img[i] = nullptr; c->tc_nrm[i] = nullptr; }
    for (int i = 0; i < 2; ++i) { XFB_ALLOC(c, c->tc_img[i], (size_t)need * 64 * 4); XFB_ALLOC(c, c->tc_nrm[i], (size_t)need * 4); }
    c->tc_cap = need;
  }
  if (c->mm_cap < need) {
    for (int i = 0; i < 2; ++i) { if (c->mm_img[i]) cudaFree(c->mm_img[i]); c->mm_img[i] = nullptr; }
    for (int i = 0; i < 2; ++i) XFB_ALLOC(c, c->mm_img[i], mm_image_bytes(need));
    c->mm_cap = need;
  }
  for (int i = 0; i < 2; ++i) if (!c->tc_nmax[i]) XFB_ALLOC(c, c->tc_nmax[i], 16);
  if (!c->tc_dbg) XFB_ALLOC(c, c->tc_dbg, 16);
  if (c->ms_cap < need) {
    for (int i = 0; i < 2; ++i) { if (c->ms_img[i]) cudaFree(c->ms_img[i]); c->ms_img[i] = nullptr; }
    for (int i = 0; i < 2; ++i) XFB_ALLOC(c, c->ms_img[i], ms_image_bytes(need));
    c->ms_cap = need;
  }
  return XFB_OK;
}

// per-pair column scratch of the mutual matcher: 64 pairs x `rows` columns, in its between-launches state
static int mm_ensure_cols(Ctx* c, int rows) {
  if (rows <= c->mm_col_rows) return XFB_OK;
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(c->mm_colg); fr(c->mm_colk); fr(c->mm_done);
  c->mm_colg = nullptr; c->mm_colk = nullptr; c->mm_done = nullptr; c->mm_col_rows = 0;
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  XFB_ALLOC(c, c->mm_colg, (size_t)64 * rows * 4);
  XFB_ALLOC(c, c->mm_colk, (size_t)64 * rows * 8);
  XFB_ALLOC(c, c->mm_done, 64 * 4);
  XFB_CUDA_OK(c, cudaMemset(c->mm_colg, 0, (size_t)64 * rows * 4));
  XFB_CUDA_OK(c, cudaMemset(c->mm_colk, 0xff, (size_t)64 * rows * 8));
  XFB_CUDA_OK(c, cudaMemset(c->mm_done, 0, 64 * 4));
  c->mm_col_rows = rows;
  return XFB_OK;
}

// A [n1,64], B [n2,64] device fp32; outputs device int32 (nullable)
static int tc_match_generic(Ctx* c, const float* dA, int n1, const float* dB, int n2, const int32_t* ga, const int32_t* gb, int init,
                            int32_t* bi, int32_t* bd, int32_t* sd, int32_t* ri, int32_t* rd) {
  int r = tc_ensure_generic(c, n1, n2);
  if (r != XFB_OK) return r;
  const int p1 = pad128(n1), p2 = pad128(n2);
  const bool grouped = ga && gb;
  if (!grouped) {
    // one streamed GEMM: row-wise best / second-best and column-wise best together (match_mutual.cu)
    if ((r = mm_ensure_cols(c, p2)) != XFB_OK) return r;
    XFB_CUDA_OK(c, launch_mm_prep(c, dA, 0, 1, nullptr, n1, p1, c->mm_img[0], 0, c->tc_nrm[0], c->tc_nmax[0]));
    XFB_CUDA_OK(c, launch_mm_prep(c, dB, 0, 1, nullptr, n2, p2, c->mm_img[1], 0, c->tc_nrm[1], c->tc_nmax[1]));
    MatchTcArgs m = {};
    m.init = init;
    m.imgA = (const float*)c->mm_img[0]; m.imgB = (const float*)c->mm_img[1]; m.nrmA = c->tc_nrm[0]; m.nrmB = c->tc_nrm[1]; m.rawA = dA; m.rawB = dB;
    m.nA_host = n1; m.nB_host = n2; m.rows_padded_A = p1; m.rows_padded_B = p2;
    m.out_stride = n1; m.out_stride_cols = n2;         // one pair: outputs are plain arrays of n1 (rows) / n2 (columns) entries
    m.best_idx = bi; m.best_dist = bd; m.second_dist = sd; m.rev_idx = ri; m.rev_dist = rd;
    m.nrm_max_A = c->tc_nmax[0]; m.nrm_max_B = c->tc_nmax[1];
    m.col_g = c->mm_colg; m.col_k = c->mm_colk; m.pair_done = c->mm_done;
    m.ms_counters = c->ms_counters;
    XFB_CUDA_OK(c, launch_match_mutual(c, m, 1, ri || rd));
    return XFB_OK;
  }
  XFB_CUDA_OK(c, launch_ms_prep(c, dA, 0, 1, nullptr, n1, p1, c->ms_img[0], 0, c->tc_nrm[0], c->tc_nmax[0]));
  XFB_CUDA_OK(c, launch_ms_prep(c, dB, 0, 1, nullptr, n2, p2, c->ms_img[1], 0, c->tc_nrm[1], c->tc_nmax[1]));
  MatchTcArgs a = {};
  a.init = init;
  a.ms_counters = c->ms_counters; a.ms_mode = c->ms_mode;
  if (n1 > 0 && (bi || bd || sd)) {
    a.imgA = (const float*)c->ms_img[0]; a.imgB = (const float*)c->ms_img[1]; a.nrmA = c->tc_nrm[0]; a.nrmB = c->tc_nrm[1]; a.rawA = dA; a.rawB = dB;
    a.gA = ga; a.gB = gb; a.nA_host = n1; a.nB_host = n2; a.rows_padded_A = p1; a.rows_padded_B = p2; a.out_stride = n1;
    a.best_idx = bi; a.best_dist = bd; a.second_dist = sd;
    a.nrm_max_B = c->tc_nmax[1];
    XFB_CUDA_OK(c, launch_match_stream(c, a, 1, grouped));
  }
  if (n2 > 0 && (ri || rd)) {   // column-wise best == row-wise best of the transposed problem (distances are symmetric, bit for bit)
    a.imgA = (const float*)c->ms_img[1]; a.imgB = (const float*)c->ms_img[0]; a.nrmA = c->tc_nrm[1]; a.nrmB = c->tc_nrm[0]; a.rawA = dB; a.rawB = dA;
    a.gA = gb; a.gB = ga; a.nA_host = n2; a.nB_host = n1; a.rows_padded_A = p2; a.rows_padded_B = p1; a.out_stride = n2;
    a.best_idx = ri; a.best_dist = rd; a.second_dist = nullptr;
    a.nrm_max_B = c->tc_nmax[0];
    XFB_CUDA_OK(c, launch_match_stream(c, a, 1, grouped));
  }
  return XFB_OK;
}

static int tc_matrix_generic(Ctx* c, const float* dA, int n1, const float* dB, int n2, int32_t* d_out, float* dbg) {
  int r = tc_ensure_generic(c, n1, n2);
  if (r != XFB_OK) return r;
  const int p1 = pad128(n1), p2 = pad128(n2);
  XFB_CUDA_OK(c, launch_match_prep(c, dA, 0, 1, nullptr, n1, p1, c->tc_img[0], 0, c->tc_nrm[0], c->tc_nmax[0]));
  XFB_CUDA_OK(c, launch_match_prep(c, dB, 0, 1, nullptr, n2, p2, c->tc_img[1], 0, c->tc_nrm[1], c->tc_nmax[1]));
  MatchTcArgs a = {};
  a.imgA = c->tc_img[0]; a.imgB = c->tc_img[1]; a.nrmA = c->tc_nrm[0]; a.nrmB = c->tc_nrm[1]; a.rawA = dA; a.rawB = dB;
  a.nA_host = n1; a.nB_host = n2; a.rows_padded_A = p1; a.rows_padded_B = p2; a.matrix = d_out; a.dbg_maxerr = dbg;
  XFB_CUDA_OK(c, launch_matrix_tc(c, a, p1 / 128));
  return XFB_OK;
}

// frames of the last extract: pairs (host) -> outputs [n_pairs][K] (device pointers, nullable)
static int tc_match_frames(Ctx* c, const int32_t* pairs, int n_pairs, int init, int32_t* o[5]) {
  const int K = c->last_topk, P = pad128(K);
  // xfb_submit runs its matcher on s_match: any other stream must not touch the shared per-frame images before that is done
  if (c->ev_match_free && c->stream != c->s_match) XFB_CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->ev_match_free, 0));
  if (!c->mm_fimg || c->mm_frows < P) {
    if (c->mm_fimg) cudaFree(c->mm_fimg);
    if (c->tc_fnrm) cudaFree(c->tc_fnrm);
    if (c->tc_fnmax) cudaFree(c->tc_fnmax);
    c->mm_fimg = nullptr; c->tc_fnrm = nullptr; c->tc_fnmax = nullptr;
    const int PM = pad128(c->max_topk);
    XFB_ALLOC(c, c->mm_fimg, (size_t)c->max_batch * mm_image_bytes(PM));
    XFB_ALLOC(c, c->tc_fnrm, (size_t)c->max_batch * PM * 4);
    XFB_ALLOC(c, c->tc_fnmax, (size_t)c->max_batch * 4);
    c->mm_frows = PM;
    c->tc_fvalid = false;
  }
  int r = mm_ensure_cols(c, pad128(c->max_topk));
  if (r != XFB_OK) return r;
  const size_t img_bytes = mm_image_bytes(P);
  if (!c->tc_fvalid) {   // operand images of every frame of the batch, once per extract
    XFB_CUDA_OK(c, launch_mm_prep(c, c->last_desc, (size_t)K * 64, c->B, c->last_nvalid, K, P, c->mm_fimg, img_bytes, c->tc_fnrm, c->tc_fnmax));
    c->tc_fvalid = true;
  }
  for (int p0 = 0; p0 < n_pairs; p0 += 64) {
    const int np = (n_pairs - p0) < 64 ? (n_pairs - p0) : 64;
    MatchTcArgs a = {};
    a.imgA = a.imgB = (const float*)c->mm_fimg; a.nrmA = a.nrmB = c->tc_fnrm; a.rawA = a.rawB = c->last_desc;
    a.nA_dev = a.nB_dev = c->last_nvalid; a.nA_host = a.nB_host = K; a.rows_padded_A = a.rows_padded_B = P;
    a.img_stride_A = a.img_stride_B = img_bytes; a.raw_stride_A = a.raw_stride_B = (size_t)K * 64;
    a.init = init; a.out_stride = K; a.out_stride_cols = K;
    a.nrm_max_A = a.nrm_max_B = c->tc_fnmax;
    a.col_g = c->mm_colg; a.col_k = c->mm_colk; a.pair_done = c->mm_done;
    a.ms_counters = c->ms_counters; a.ms_mode = c->ms_mode;
    for (int p = 0; p < np; ++p) { a.pairs[2 * p] = pairs[2 * (p0 + p)]; a.pairs[2 * p + 1] = pairs[2 * (p0 + p) + 1]; }
    auto at = [&](int i) { return o[i] ? o[i] + (size_t)p0 * K : nullptr; };
    a.best_idx = at(0); a.best_dist = at(1); a.second_dist = at(2); a.rev_idx = at(3); a.rev_dist = at(4);
    // ONE launch per 64 pairs: row-wise best / second-best and column-wise best from one streamed GEMM per pair (match_mutual.cu)
    XFB_CUDA_OK(c, launch_match_mutual(c, a, np, o[3] || o[4]));
  }
  return XFB_OK;
}

// ---- the forward pipeline (XFeatModel::forward, src/XFeat.cc:135-173, then :273-316) -----------------
static int run_dense(Ctx* c, const uint8_t* d_gray, size_t frame_stride, int stride) {
  XFB_CUDA_OK(c, launch_prep(c, d_gray, frame_stride, stride));
  static const int order1[] = {L_B1_0, L_B1_1, L_B1_2, L_B1_3, L_B2_0, L_B2_1, L_B3_0, L_B3_1, L_B3_2,
                               L_B4_0, L_B4_1, L_B4_2, L_B5_0, L_B5_1, L_B5_2, L_B5_3};
  auto conv = [&](int L) {
    if (c->force_simt) return launch_conv_layer(c, L);                 // generic FP32 SIMT kernels (A/B reference)
    if (conv_tc_handles(L)) return c->conv_tc_version == 2 ? launch_conv_tc2_layer(c, L) : launch_conv_tc_layer(c, L);   // tcgen05 implicit GEMM (Cin >= 24)
    if (conv_small_handles(L)) return launch_conv_small_layer(c, L);   // block1: bandwidth-shaped SIMT
    return launch_conv_layer(c, L);
  };
  for (int L : order1) XFB_CUDA_OK(c, conv(L));
  XFB_CUDA_OK(c, launch_pyramid(c));
  static const int order2[] = {L_F_0, L_F_1, L_F_2, L_HM_0, L_HM_1};
  for (int L : order2) XFB_CUDA_OK(c, conv(L));
  XFB_CUDA_OK(c, launch_heatmap_out(c));
  static const int order3[] = {L_KP_0, L_KP_1, L_KP_2};
  for (int L : order3) XFB_CUDA_OK(c, conv(L));
  if (c->force_simt) XFB_CUDA_OK(c, launch_keypoint_out(c));
  else XFB_CUDA_OK(c, c->conv_tc_version == 2 ? launch_conv_tc2_layer(c, L_KP_3) : launch_conv_tc_layer(c, L_KP_3));
  return XFB_OK;
}

static int check_extract_args(Ctx* c, int batch, int h, int w, int stride, int topk) {
  if (!c) return XFB_ERR_ARG;
  if (h <= 0 || w <= 0 || batch <= 0) { c->err = "empty image"; return XFB_ERR_EMPTY; }
  if (h < 32 || w < 32 || h > c->max_h || w > c->max_w || stride < w || batch > c->max_batch || topk < 1 || topk > c->max_topk) {
    c->err = "extract: size/batch/topk outside the limits given to xfb_create (image must be >= 32x32)";
    return XFB_ERR_ARG;
  }
  if ((h / 32) * (w / 32) < 2) {
    // block5 would normalise ONE value per channel: libtorch's train-mode batch_norm throws here
    // ("Expected more than 1 value per channel when training"), i.e. the reference cannot process such frames
    c->err = "extract: image too small -- the 1/32-resolution map has a single pixel (the reference's train-mode BatchNorm throws)";
    return XFB_ERR_ARG;
  }
  return XFB_OK;
}

static int extract_device(Ctx* c, const uint8_t* d_gray, int batch, size_t frame_stride, int h, int w, int stride, int topk,
                          float nms_thr, int32_t* d_nvalid, float* d_xy, float* d_score, float* d_desc) {
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  c->B = batch; c->in_h = h; c->in_w = w;
  c->H = (h / 32) * 32; c->W = (w / 32) * 32;
  int r = run_dense(c, d_gray, frame_stride, stride);
  if (r != XFB_OK) return r;
  XFB_CUDA_OK(c, launch_post(c, topk, nms_thr, d_nvalid, d_xy, d_score, d_desc));
  c->last_topk = topk; c->last_nvalid = d_nvalid; c->last_desc = d_desc; c->tc_fvalid = false;
  return XFB_OK;
}

}  // namespace xfb

using namespace xfb;

extern "C" {

int xfb_create(xfb_ctx** out, const void* weights_blob, size_t n, int device, int max_h, int max_w, int max_batch, int max_topk) {
  if (!out || !weights_blob || max_h < 32 || max_w < 32 || max_batch < 1 || max_topk < 1 || max_topk > 8192) {
    set_global_error("xfb_create: bad argument");
    return XFB_ERR_ARG;
  }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) {
    set_global_error(std::string("xfb_create: no usable CUDA device (there is no CPU fallback): ") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "device ordinal out of range"));
    return XFB_ERR_CUDA;
  }
  xfb_ctx* c = new xfb_ctx();
  c->device = device; c->max_h = max_h; c->max_w = max_w; c->max_batch = max_batch; c->max_topk = max_topk;
  int r = XFB_OK;
  do {
    if ((e = cudaSetDevice(device)) != cudaSuccess) { c->err = cudaGetErrorString(e); r = XFB_ERR_CUDA; break; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { c->err = cudaGetErrorString(e); r = XFB_ERR_CUDA; break; }
    if (prop.major != 10) {
      c->err = "libxfeat_b200 is built for sm_100a only; device reports sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
      r = XFB_ERR_CUDA;
      break;
    }
    c->num_sms = prop.multiProcessorCount;
    if (const char* bf = getenv("XFB_B1_FUSE")) c->b1_fuse = atoi(bf) != 0;
    if (const char* pd = getenv("XFB_PDL")) c->pdl = atoi(pd) != 0;
    if (const char* cv = getenv("XFB_CONV_TC")) c->conv_tc_version = (atoi(cv) == 1) ? 1 : 2;   // A/B: 1 = the round-1 one-tile-per-CTA 3xTF32 kernels
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) { c->err = cudaGetErrorString(e); r = XFB_ERR_CUDA; break; }
    c->stream = c->own_stream;
    if (const char* md = getenv("XFB_MS_DEBUG")) {   // profiling aid (match_stream.cu): cycle counters of CTA (0,0) printed at xfb_destroy
      // bit 64 also counts survivors; the other bits switch kernel stages OFF for timing ablations (results are then WRONG), so they
      // are honoured only together with XFB_MS_DEBUG_ABLATE=1
      c->ms_mode = atoi(md) & (getenv("XFB_MS_DEBUG_ABLATE") ? ~0 : 64);
      if ((e = cudaMalloc(&c->ms_counters, 512)) != cudaSuccess || (e = cudaMemset(c->ms_counters, 0, 512)) != cudaSuccess) { c->err = cudaGetErrorString(e); r = XFB_ERR_CUDA; break; }
    }
    if (getenv("XFB_T2_DEBUG")) {   // profiling aid (conv_tc2.cu): per-role cycle counters of CTA 0, printed per layer at xfb_destroy
      if ((e = cudaMalloc(&c->t2_counters, L_NUM * 32 * 8)) != cudaSuccess || (e = cudaMemset(c->t2_counters, 0, L_NUM * 32 * 8)) != cudaSuccess) { c->err = cudaGetErrorString(e); r = XFB_ERR_CUDA; break; }
    }
    if ((r = load_weights(c, static_cast<const uint8_t*>(weights_blob), n)) != XFB_OK) break;
    if ((r = alloc_buffers(c)) != XFB_OK) break;
  } while (0);
  if (r != XFB_OK) {
    set_global_error(c->err);
    xfb_destroy(c);
    return r;
  }
  *out = c;
  return XFB_OK;
}

void xfb_destroy(xfb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->own_stream) cudaStreamSynchronize(c->own_stream);
  auto fr = [](void* p) { if (p) cudaFree(p); };
  for (int L = 0; L < L_NUM; ++L) { fr(c->w[L]); fr(c->bias[L]); fr(c->act[L]); fr(c->wimg[L]); fr(c->wimg2[L]); }
  for (int L = 0; L < L_NUM_BN; ++L) { fr(c->bn[L].mean); fr(c->bn[L].rstd); }
  fr(c->d_gray); fr(c->xraw); fr(c->xn); fr(c->avg4); fr(c->pyr); fr(c->k1h); fr(c->in_mean); fr(c->in_rstd); fr(c->part);
  fr(c->ticket); fr(c->cand); fr(c->cand_count); fr(c->cand_count_last); fr(c->o_nvalid); fr(c->o_xy); fr(c->o_score); fr(c->o_desc);
  fr(c->m_a); fr(c->m_b); fr(c->m_ga); fr(c->m_gb); fr(c->m_matrix);
  for (int i = 0; i < 5; ++i) { fr(c->m_out[i]); fr(c->m_pairs_out[i]); }
  for (int i = 0; i < 2; ++i) { fr(c->tc_img[i]); fr(c->tc_nrm[i]); }
  if (c->t2_counters) {
    std::vector<unsigned long long> h(L_NUM * 32);
    cudaMemcpy(h.data(), c->t2_counters, h.size() * 8, cudaMemcpyDeviceToHost);
    for (int L = 0; L < L_NUM; ++L) {
      const unsigned long long* q = &h[(size_t)L * 32];
      if (!q[0]) continue;
      const double n = (double)q[0], tl = q[16] ? (double)q[16] : 1.0, st = q[5] ? (double)q[5] : 1.0;
      fprintf(stderr, "[xfb t2] %-16s launches %llu | CTA0 cycles/launch %.0f, tiles/launch %.1f | producer per stage: wait_free %.0f work %.0f fence+arrive %.0f | "
              "mma per tile: wait_in %.0f wait_acce %.0f issue %.0f wait_w(total/launch) %.0f | epilogue per tile: wait_accf %.0f tmem ld %.0f "
              "staging stores + TMA issue %.0f column sums + partials %.0f staging-tile barrier %.0f publish + loop %.0f\n",
              kLayers[L].ref_name, q[0], q[1] / n, tl / n, q[2] / st, q[3] / st, q[4] / st, q[6] / tl, q[7] / tl, q[8] / tl, q[9] / n, q[10] / tl, q[11] / tl,
              q[12] / tl, q[13] / tl, q[14] / tl, q[15] / tl);
    }
    cudaFree(c->t2_counters);
  }
  if (c->ms_counters) {
    unsigned long long h[64] = {};
    cudaMemcpy(h, c->ms_counters, 512, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[xfb] match_stream counters: pushes %llu, verified %llu, overflow %llu\n", h[0], h[1], h[2]);
    const double n = h[13] ? (double)h[13] : 1.0;
    fprintf(stderr, "[xfb] CTA(0,0) cycles per launch: total %.0f | mma thread: wait A %.0f, wait full %.0f, wait acc-empty %.0f, loop %.0f | loader wait empty %.0f | "
            "epilogue warp 2: wait acc-full (pass 1) %.0f, pass-1 end @%.0f, pass-2 end @%.0f, drain end @%.0f\n",
            h[12] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[7] / n, h[8] / n, h[9] / n, h[10] / n, h[11] / n);
    if (h[30]) {
      const double m = (double)h[30];
      fprintf(stderr, "[xfb] match_mutual CTA(7,3) cycles per launch: total %.0f | MMA lane loop end @%.0f | row warp: stream+bound end @%.0f, drain end @%.0f | "
              "column warp: stream end @%.0f, drain end @%.0f || all CTAs per launch: row entries %.0f verified %.0f, column entries %.0f verified %.0f\n",
              h[20] / m, h[25] / m, h[21] / m, h[22] / m, h[23] / m, h[24] / m, h[26] / m, h[27] / m, h[28] / m, h[29] / m);
      fprintf(stderr, "[xfb]   MMA lane per launch: wait column block %.0f, wait column accumulator free %.0f, wait row accumulator free %.0f, issue %.0f | "
              "row warp 0 (half of the blocks): wait accumulator %.0f, work %.0f | column warp 0: wait accumulator %.0f, work %.0f\n",
              h[52] / m, h[53] / m, h[54] / m, h[55] / m, h[56] / m, h[57] / m, h[58] / m, h[59] / m);
      const char* nm[4] = {"row overflow", "row final", "column overflow", "column final"};
      for (int k = 0; k < 4; ++k) {
        const unsigned long long* d = h + 32 + 5 * k;
        fprintf(stderr, "[xfb]   %s drains of ONE warp of that CTA, per launch: calls %.2f, entries in %.0f, kept %.0f, filter cycles %.0f, verify cycles %.0f\n",
                nm[k], d[0] / m, d[1] / m, d[2] / m, d[3] / m, d[4] / m);
      }
    }
    fprintf(stderr, "[xfb] CTA(0,0) mma thread: cycles in tcgen05.mma issue %.0f, in tcgen05.commit %.0f, in tcgen05.fence %.0f, loop tail %.0f\n", h[14] / n, h[15] / n, h[16] / n, h[17] / n);
  }
  fr(c->ms_img[0]); fr(c->ms_img[1]); fr(c->ms_fimg); fr(c->mm_img[0]); fr(c->mm_img[1]); fr(c->mm_fimg); fr(c->mm_colg); fr(c->mm_colk); fr(c->mm_done); fr(c->ms_counters); fr(c->p_idx); fr(c->p_out); fr(c->g_buf); fr(c->v_desc); fr(c->v_start); fr(c->v_child); fr(c->v_out);
  fr(c->tc_fnrm); fr(c->tc_pairs); fr(c->tc_dbg); fr(c->tc_fnmax); fr(c->tc_nmax[0]); fr(c->tc_nmax[1]);
  for (auto& s : c->slots) {
    fr(s.d_gray); fr(s.nvalid); fr(s.xy); fr(s.score); fr(s.desc);
    for (int i = 0; i < 5; ++i) fr(s.m[i]);
    if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
    if (s.ev_comp) cudaEventDestroy(s.ev_comp);
    if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
    if (s.ev_ext) cudaEventDestroy(s.ev_ext);
  }
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  if (c->s_match) cudaStreamDestroy(c->s_match);
  if (c->ev_match_free) cudaEventDestroy(c->ev_match_free);
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

const char* xfb_last_error(const xfb_ctx* c) {
  if (c) return c->err.c_str();
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_err_mu);
  copy = g_err;
  return copy.c_str();
}

int xfb_set_stream(xfb_ctx* c, void* cuda_stream) {
  if (!c) return XFB_ERR_ARG;
  c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return XFB_OK;
}

int xfb_extract_batch_device(xfb_ctx* c, const uint8_t* d_gray, int batch, size_t frame_stride, int h, int w, int stride, int topk,
                             float nms_thr, int32_t* d_n_valid, float* d_kpt_xy, float* d_score, float* d_desc) {
  int r = check_extract_args(c, batch, h, w, stride, topk);
  if (r != XFB_OK) return r;
  if (!d_gray || !d_n_valid || !d_kpt_xy || !d_score || !d_desc) { c->err = "extract: null pointer"; return XFB_ERR_ARG; }
  return extract_device(c, d_gray, batch, frame_stride, h, w, stride, topk, nms_thr, d_n_valid, d_kpt_xy, d_score, d_desc);
}

int xfb_extract_batch(xfb_ctx* c, const uint8_t* gray, int batch, size_t frame_stride, int h, int w, int stride, int topk,
                      float nms_thr, int32_t* n_valid, float* kpt_xy, float* score, float* desc) {
  int r = check_extract_args(c, batch, h, w, stride, topk);
  if (r != XFB_OK) return r;
  if (!gray || !n_valid || !kpt_xy || !score || !desc) { c->err = "extract: null pointer"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  // host -> device: rows are packed to `w` bytes on the device
  const size_t dev_fs = (size_t)h * w;
  for (int b = 0; b < batch; ++b)
    XFB_CUDA_OK(c, cudaMemcpy2DAsync(c->d_gray + b * dev_fs, w, gray + b * frame_stride, stride, w, h, cudaMemcpyHostToDevice, c->stream));
  r = extract_device(c, c->d_gray, batch, dev_fs, h, w, w, topk, nms_thr, c->o_nvalid, c->o_xy, c->o_score, c->o_desc);
  if (r != XFB_OK) return r;
  const size_t K = topk;
  XFB_CUDA_OK(c, cudaMemcpyAsync(n_valid, c->o_nvalid, (size_t)batch * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(kpt_xy, c->o_xy, batch * K * 2 * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(score, c->o_score, batch * K * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(desc, c->o_desc, batch * K * 64 * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return XFB_OK;
}

int xfb_extract(xfb_ctx* c, const uint8_t* gray, int h, int w, int stride, int topk, float nms_thr, int32_t* n_valid, float* kpt_xy,
                float* score, float* desc) {
  return xfb_extract_batch(c, gray, 1, (size_t)h * (size_t)(stride > 0 ? stride : 0), h, w, stride, topk, nms_thr, n_valid, kpt_xy, score, desc);
}

int xfb_distance_matrix_device(xfb_ctx* c, const float* d_A, int n1, const float* d_B, int n2, int32_t* d_out) {
  if (!c) return XFB_ERR_ARG;
  if (n1 < 0 || n2 < 0 || (n1 && !d_A) || (n2 && !d_B) || (n1 && n2 && !d_out)) { c->err = "distance_matrix: bad argument"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  if (n1 == 0 || n2 == 0) return XFB_OK;
  return tc_matrix_generic(c, d_A, n1, d_B, n2, d_out, nullptr);
}

int xfb_distance_matrix(xfb_ctx* c, const float* A, int n1, const float* B, int n2, int32_t* out) {
  if (!c) return XFB_ERR_ARG;
  if (n1 < 0 || n2 < 0 || (n1 && !A) || (n2 && !B) || (n1 && n2 && !out)) { c->err = "distance_matrix: bad argument"; return XFB_ERR_ARG; }
  if (n1 == 0 || n2 == 0) return XFB_OK;
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  int r = ensure_match_scratch(c, n1, n2, true);
  if (r != XFB_OK) return r;
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_a, A, (size_t)n1 * 256, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_b, B, (size_t)n2 * 256, cudaMemcpyHostToDevice, c->stream));
  r = tc_matrix_generic(c, c->m_a, n1, c->m_b, n2, c->m_matrix, nullptr);
  if (r != XFB_OK) return r;
  XFB_CUDA_OK(c, cudaMemcpyAsync(out, c->m_matrix, (size_t)n1 * n2 * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return XFB_OK;
}

// ---- per-keypoint frame geometry (geom.cu) ---------------------------------------------------------------------
int xfb_image_bounds(xfb_camera* cam, int w, int h) {
  if (!cam || w <= 0 || h <= 0 || cam->fx == 0.f || cam->fy == 0.f) return XFB_ERR_ARG;
  image_bounds_host(cam, w, h);
  return XFB_OK;
}

int xfb_keypoint_geometry_device(xfb_ctx* c, const float* d_xy, int n, const float* d_depth, int h, int w, int depth_stride, const xfb_camera* cam,
                                 float* d_un, float* d_kd, float* d_ur, int32_t* d_cell) {
  if (!c) return XFB_ERR_ARG;
  if (n < 0 || !cam || (n && !d_xy) || (d_depth && (h <= 0 || w <= 0 || depth_stride < w)) || cam->fx == 0.f || cam->fy == 0.f ||
      (d_cell && (!(cam->max_x > cam->min_x) || !(cam->max_y > cam->min_y)))) { c->err = "keypoint_geometry: bad argument"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  XFB_CUDA_OK(c, launch_keypoint_geometry(c, d_xy, n, d_depth, h, w, depth_stride, *cam, d_un, d_kd, d_ur, d_cell));
  return XFB_OK;
}

int xfb_keypoint_geometry(xfb_ctx* c, const float* xy, int n, const float* depth, int h, int w, int depth_stride, const xfb_camera* cam, float* un_xy,
                          float* kp_depth, float* uright, int32_t* cell) {
  if (!c) return XFB_ERR_ARG;
  if (n < 0 || !cam || (n && !xy) || (depth && (h <= 0 || w <= 0 || depth_stride < w)) || cam->fx == 0.f || cam->fy == 0.f ||
      (cell && (!(cam->max_x > cam->min_x) || !(cam->max_y > cam->min_y)))) { c->err = "keypoint_geometry: bad argument"; return XFB_ERR_ARG; }
  if (n == 0) return XFB_OK;
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  const size_t img = depth ? (size_t)h * w * 4 : 0, pts = (size_t)n * 4;
  const size_t need = img + pts * 7;   // depth | xy (2) | un (2) | depth out | uright | cell
  if (need > c->g_cap) {
    if (c->g_buf) cudaFree(c->g_buf);
    c->g_buf = nullptr; c->g_cap = 0;
    XFB_ALLOC(c, c->g_buf, need + (1 << 20));
    c->g_cap = need + (1 << 20);
  }
  float* d_depth = depth ? c->g_buf : nullptr;
  float* d_xy = reinterpret_cast<float*>(reinterpret_cast<char*>(c->g_buf) + img);
  float* d_un = d_xy + 2 * (size_t)n; float* d_kd = d_un + 2 * (size_t)n; float* d_ur = d_kd + n;
  int32_t* d_cell = reinterpret_cast<int32_t*>(d_ur + n);
  if (depth) XFB_CUDA_OK(c, cudaMemcpy2DAsync(d_depth, (size_t)w * 4, depth, (size_t)depth_stride * 4, (size_t)w * 4, h, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(d_xy, xy, pts * 2, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, launch_keypoint_geometry(c, d_xy, n, d_depth, h, w, w, *cam, d_un, d_kd, d_ur, d_cell));
  if (un_xy) XFB_CUDA_OK(c, cudaMemcpyAsync(un_xy, d_un, pts * 2, cudaMemcpyDeviceToHost, c->stream));
  if (kp_depth) XFB_CUDA_OK(c, cudaMemcpyAsync(kp_depth, d_kd, pts, cudaMemcpyDeviceToHost, c->stream));
  if (uright) XFB_CUDA_OK(c, cudaMemcpyAsync(uright, d_ur, pts, cudaMemcpyDeviceToHost, c->stream));
  if (cell) XFB_CUDA_OK(c, cudaMemcpyAsync(cell, d_cell, pts, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return XFB_OK;
}

// ---- vocabulary tree walk (bow.cu) -----------------------------------------------------------------------------
int xfb_vocab_load(xfb_ctx* c, const uint8_t* node_desc, const int32_t* child_start, const int32_t* child_index, int n_nodes, int n_children, int L) {
  if (!c) return XFB_ERR_ARG;
  if (!node_desc || !child_start || !child_index || n_nodes < 2 || n_children != n_nodes - 1 || L < 1 || child_start[0] != 0 ||
      child_start[n_nodes] != n_children || child_start[1] <= 0) { c->err = "vocab_load: bad argument"; return XFB_ERR_ARG; }
  for (int i = 0; i < n_nodes; ++i) if (child_start[i + 1] < child_start[i]) { c->err = "vocab_load: child_start is not monotone"; return XFB_ERR_ARG; }
  for (int k = 0; k < n_children; ++k) if (child_index[k] <= 0 || child_index[k] >= n_nodes) { c->err = "vocab_load: child index out of range"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  auto fr = [](void* p) { if (p) cudaFree(p); };
  fr(c->v_desc); fr(c->v_start); fr(c->v_child);
  c->v_desc = nullptr; c->v_start = nullptr; c->v_child = nullptr; c->v_nodes = 0;
  XFB_ALLOC(c, c->v_desc, (size_t)n_nodes * 32);
  XFB_ALLOC(c, c->v_start, (size_t)(n_nodes + 1) * 4);
  XFB_ALLOC(c, c->v_child, (size_t)n_children * 4);
  XFB_CUDA_OK(c, cudaMemcpy(c->v_desc, node_desc, (size_t)n_nodes * 32, cudaMemcpyHostToDevice));
  XFB_CUDA_OK(c, cudaMemcpy(c->v_start, child_start, (size_t)(n_nodes + 1) * 4, cudaMemcpyHostToDevice));
  XFB_CUDA_OK(c, cudaMemcpy(c->v_child, child_index, (size_t)n_children * 4, cudaMemcpyHostToDevice));
  c->v_nodes = n_nodes; c->v_L = L;
  return XFB_OK;
}

int xfb_bow_transform_device(xfb_ctx* c, const float* d_desc, int n, int levelsup, int32_t* d_leaf, int32_t* d_nid) {
  if (!c) return XFB_ERR_ARG;
  if (!c->v_nodes) { c->err = "bow_transform: no vocabulary loaded (xfb_vocab_load)"; return XFB_ERR_ARG; }
  if (n < 0 || (n && (!d_desc || !d_leaf || !d_nid))) { c->err = "bow_transform: bad argument"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  XFB_CUDA_OK(c, launch_bow_transform(c, d_desc, 0, 1, nullptr, n, levelsup, d_leaf, d_nid, n));
  return XFB_OK;
}

int xfb_bow_transform(xfb_ctx* c, const float* desc, int n, int levelsup, int32_t* leaf, int32_t* nid) {
  if (!c) return XFB_ERR_ARG;
  if (!c->v_nodes) { c->err = "bow_transform: no vocabulary loaded (xfb_vocab_load)"; return XFB_ERR_ARG; }
  if (n < 0 || (n && (!desc || !leaf || !nid))) { c->err = "bow_transform: bad argument"; return XFB_ERR_ARG; }
  if (n == 0) return XFB_OK;
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  int r = ensure_match_scratch(c, n, 0, false);
  if (r != XFB_OK) return r;
  if (n > c->v_out_cap) {
    if (c->v_out) cudaFree(c->v_out);
    c->v_out = nullptr; c->v_out_cap = 0;
    XFB_ALLOC(c, c->v_out, (size_t)c->m_cap * 2 * 4);
    c->v_out_cap = c->m_cap;
  }
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_a, desc, (size_t)n * 64 * 4, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, launch_bow_transform(c, c->m_a, 0, 1, nullptr, n, levelsup, c->v_out, c->v_out + c->v_out_cap, n));
  XFB_CUDA_OK(c, cudaMemcpyAsync(leaf, c->v_out, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(nid, c->v_out + c->v_out_cap, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return XFB_OK;
}

int xfb_bow_transform_frames_device(xfb_ctx* c, int levelsup, int32_t* d_leaf, int32_t* d_nid) {
  if (!c) return XFB_ERR_ARG;
  if (!c->v_nodes) { c->err = "bow_transform: no vocabulary loaded (xfb_vocab_load)"; return XFB_ERR_ARG; }
  if (!c->last_desc || !d_leaf || !d_nid) { c->err = "bow_transform_frames: bad argument / no extract result"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  const int K = c->last_topk;
  XFB_CUDA_OK(c, launch_bow_transform(c, c->last_desc, (size_t)K * 64, c->B, c->last_nvalid, K, levelsup, d_leaf, d_nid, K));
  return XFB_OK;
}

int xfb_distance_pairs_device(xfb_ctx* c, const float* d_A, int n1, const float* d_B, int n2, const int32_t* d_ia, const int32_t* d_ib, int n_pairs,
                               int32_t* d_out) {
  if (!c) return XFB_ERR_ARG;
  if (n1 < 0 || n2 < 0 || n_pairs < 0 || (n_pairs && (!d_A || !d_B || !d_ia || !d_ib || !d_out))) { c->err = "distance_pairs: bad argument"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  XFB_CUDA_OK(c, launch_distance_pairs(c, d_A, n1, d_B, n2, d_ia, d_ib, n_pairs, d_out));
  return XFB_OK;
}

int xfb_distance_pairs(xfb_ctx* c, const float* A, int n1, const float* B, int n2, const int32_t* idx_a, const int32_t* idx_b, int n_pairs, int32_t* out) {
  if (!c) return XFB_ERR_ARG;
  if (n1 < 0 || n2 < 0 || n_pairs < 0 || (n_pairs && (!A || !B || !idx_a || !idx_b || !out))) { c->err = "distance_pairs: bad argument"; return XFB_ERR_ARG; }
  if (n_pairs == 0) return XFB_OK;
  for (int p = 0; p < n_pairs; ++p)
    if (idx_a[p] < 0 || idx_a[p] >= n1 || idx_b[p] < 0 || idx_b[p] >= n2) { c->err = "distance_pairs: index out of range"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  int r = ensure_match_scratch(c, n1, n2, false);
  if (r != XFB_OK) return r;
  if (n_pairs > c->p_cap) {
    if (c->p_idx) cudaFree(c->p_idx);
    if (c->p_out) cudaFree(c->p_out);
    c->p_idx = nullptr; c->p_out = nullptr; c->p_cap = 0;
    const size_t cap = (size_t)(n_pairs + 4095) / 4096 * 4096;
    XFB_ALLOC(c, c->p_idx, cap * 2 * 4);
    XFB_ALLOC(c, c->p_out, cap * 4);
    c->p_cap = (int)cap;
  }
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_a, A, (size_t)n1 * 64 * 4, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_b, B, (size_t)n2 * 64 * 4, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->p_idx, idx_a, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->p_idx + c->p_cap, idx_b, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, launch_distance_pairs(c, c->m_a, n1, c->m_b, n2, c->p_idx, c->p_idx + c->p_cap, n_pairs, c->p_out));
  XFB_CUDA_OK(c, cudaMemcpyAsync(out, c->p_out, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return XFB_OK;
}

int xfb_match_device(xfb_ctx* c, const float* d_A, int n1, const float* d_B, int n2, const int32_t* d_ga, const int32_t* d_gb,
                     int init_dist, int32_t* bi, int32_t* bd, int32_t* sd, int32_t* ri, int32_t* rd) {
  if (!c) return XFB_ERR_ARG;
  if (n1 < 0 || n2 < 0 || (n1 && !d_A) || (n2 && !d_B) || ((d_ga == nullptr) != (d_gb == nullptr))) { c->err = "match: bad argument"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  int r = ensure_match_scratch(c, n1, n2, false);
  if (r != XFB_OK) return r;
  return tc_match_generic(c, d_A, n1, d_B, n2, d_ga, d_gb, init_dist, bi, bd, sd, ri, rd);
}

int xfb_match(xfb_ctx* c, const float* A, int n1, const float* B, int n2, const int32_t* ga, const int32_t* gb, int init_dist,
              int32_t* best_idx, int32_t* best_dist, int32_t* second_dist, int32_t* best_idx_rev, int32_t* best_dist_rev) {
  if (!c) return XFB_ERR_ARG;
  if (n1 < 0 || n2 < 0 || (n1 && !A) || (n2 && !B) || ((ga == nullptr) != (gb == nullptr))) { c->err = "match: bad argument"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  int r = ensure_match_scratch(c, n1, n2, false);
  if (r != XFB_OK) return r;
  if (n1) XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_a, A, (size_t)n1 * 256, cudaMemcpyHostToDevice, c->stream));
  if (n2) XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_b, B, (size_t)n2 * 256, cudaMemcpyHostToDevice, c->stream));
  if (ga) {
    if (n1) XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_ga, ga, (size_t)n1 * 4, cudaMemcpyHostToDevice, c->stream));
    if (n2) XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_gb, gb, (size_t)n2 * 4, cudaMemcpyHostToDevice, c->stream));
  }
  r = tc_match_generic(c, c->m_a, n1, c->m_b, n2, ga ? c->m_ga : nullptr, ga ? c->m_gb : nullptr, init_dist, best_idx ? c->m_out[0] : nullptr,
                       best_dist ? c->m_out[1] : nullptr, second_dist ? c->m_out[2] : nullptr, best_idx_rev ? c->m_out[3] : nullptr,
                       best_dist_rev ? c->m_out[4] : nullptr);
  if (r != XFB_OK) return r;
  int32_t* host[5] = {best_idx, best_dist, second_dist, best_idx_rev, best_dist_rev};
  for (int i = 0; i < 5; ++i) {
    const int cnt = i < 3 ? n1 : n2;
    if (host[i] && cnt) XFB_CUDA_OK(c, cudaMemcpyAsync(host[i], c->m_out[i], (size_t)cnt * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return XFB_OK;
}

static int match_pairs_impl(xfb_ctx* c, const int32_t* pairs, int n_pairs, int init_dist, int32_t* o[5], bool to_host) {
  if (!c) return XFB_ERR_ARG;
  if (!pairs || n_pairs < 0 || !c->last_desc) { c->err = "match_frame_pairs: bad argument / no extract result"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  const int K = c->last_topk;
  for (int p = 0; p < n_pairs; ++p) {
    const int fa = pairs[2 * p], fb = pairs[2 * p + 1];
    if (fa < 0 || fb < 0 || fa >= c->B || fb >= c->B) { c->err = "match_frame_pairs: frame index out of range"; return XFB_ERR_ARG; }
  }
  if (to_host && n_pairs > c->m_pairs_cap) {
    for (int i = 0; i < 5; ++i) { if (c->m_pairs_out[i]) cudaFree(c->m_pairs_out[i]); c->m_pairs_out[i] = nullptr; }
    for (int i = 0; i < 5; ++i) XFB_ALLOC(c, c->m_pairs_out[i], (size_t)n_pairs * c->max_topk * 4);
    c->m_pairs_cap = n_pairs;
  }
  int32_t* d[5];
  for (int i = 0; i < 5; ++i) d[i] = o[i] ? (to_host ? c->m_pairs_out[i] : o[i]) : nullptr;
  int r = tc_match_frames(c, pairs, n_pairs, init_dist, d);
  if (r != XFB_OK) return r;
  if (to_host) {
    for (int i = 0; i < 5; ++i)
      if (o[i] && n_pairs) XFB_CUDA_OK(c, cudaMemcpyAsync(o[i], c->m_pairs_out[i], (size_t)n_pairs * K * 4, cudaMemcpyDeviceToHost, c->stream));
    XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  }
  return XFB_OK;
}

int xfb_match_frames(xfb_ctx* c, int fa, int fb, int init_dist, int32_t* best_idx, int32_t* best_dist, int32_t* second_dist,
                     int32_t* best_idx_rev, int32_t* best_dist_rev) {
  const int32_t pair[2] = {fa, fb};
  int32_t* o[5] = {best_idx, best_dist, second_dist, best_idx_rev, best_dist_rev};
  return match_pairs_impl(c, pair, 1, init_dist, o, true);
}

int xfb_match_frame_pairs(xfb_ctx* c, const int32_t* pairs, int n_pairs, int init_dist, int32_t* bi, int32_t* bd, int32_t* sd, int32_t* ri,
                          int32_t* rd) {
  int32_t* o[5] = {bi, bd, sd, ri, rd};
  return match_pairs_impl(c, pairs, n_pairs, init_dist, o, true);
}
int xfb_match_frame_pairs_device(xfb_ctx* c, const int32_t* pairs, int n_pairs, int init_dist, int32_t* bi, int32_t* bd, int32_t* sd,
                                 int32_t* ri, int32_t* rd) {
  int32_t* o[5] = {bi, bd, sd, ri, rd};
  return match_pairs_impl(c, pairs, n_pairs, init_dist, o, false);
}

static int slot_ensure(xfb_ctx* c, Ctx::Slot& s, int n_pairs) {
  if (!s.d_gray) {
    const size_t B = c->max_batch, K = c->max_topk;
    XFB_ALLOC(c, s.d_gray, B * (size_t)c->max_h * c->max_w);
    XFB_ALLOC(c, s.nvalid, B * 4);
    XFB_ALLOC(c, s.xy, B * K * 2 * 4);
    XFB_ALLOC(c, s.score, B * K * 4);
    XFB_ALLOC(c, s.desc, B * K * 64 * 4);
    XFB_CUDA_OK(c, cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
    XFB_CUDA_OK(c, cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming));
    XFB_CUDA_OK(c, cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming));
    XFB_CUDA_OK(c, cudaEventCreateWithFlags(&s.ev_ext, cudaEventDisableTiming));
  }
  if (n_pairs > s.m_cap) {
    for (int i = 0; i < 5; ++i) { if (s.m[i]) cudaFree(s.m[i]); s.m[i] = nullptr; }
    for (int i = 0; i < 5; ++i) XFB_ALLOC(c, s.m[i], (size_t)n_pairs * c->max_topk * 4);
    s.m_cap = n_pairs;
  }
  if (!c->s_h2d) XFB_CUDA_OK(c, cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
  if (!c->s_d2h) XFB_CUDA_OK(c, cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
  if (!c->s_match) {
    XFB_CUDA_OK(c, cudaStreamCreateWithFlags(&c->s_match, cudaStreamNonBlocking));
    XFB_CUDA_OK(c, cudaEventCreateWithFlags(&c->ev_match_free, cudaEventDisableTiming));
    XFB_CUDA_OK(c, cudaEventRecord(c->ev_match_free, c->s_match));
  }
  return XFB_OK;
}

int xfb_wait(xfb_ctx* c, int slot) {
  if (!c || slot < 0 || slot > 1) return XFB_ERR_ARG;
  Ctx::Slot& s = c->slots[slot];
  if (s.pending) {
    XFB_CUDA_OK(c, cudaEventSynchronize(s.ev_d2h));
    s.pending = false;
  }
  return XFB_OK;
}

int xfb_submit(xfb_ctx* c, int slot, const uint8_t* gray, int batch, size_t frame_stride, int h, int w, int stride, int topk, float nms_thr,
               int32_t* n_valid, float* kpt_xy, float* score, float* desc, const int32_t* pairs, int n_pairs, int init_dist, int32_t* bi,
               int32_t* bd, int32_t* sd, int32_t* ri, int32_t* rd) {
  int r = check_extract_args(c, batch, h, w, stride, topk);
  if (r != XFB_OK) return r;
  if (slot < 0 || slot > 1 || !gray || !n_valid || !kpt_xy || !score || !desc || n_pairs < 0 || (n_pairs && !pairs)) {
    c->err = "submit: bad argument";
    return XFB_ERR_ARG;
  }
  for (int p = 0; p < n_pairs; ++p)
    if (pairs[2 * p] < 0 || pairs[2 * p + 1] < 0 || pairs[2 * p] >= batch || pairs[2 * p + 1] >= batch) { c->err = "submit: pair index out of range"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  if ((r = xfb_wait(c, slot)) != XFB_OK) return r;             // the slot's previous outputs must have been delivered
  Ctx::Slot& s = c->slots[slot];
  if ((r = slot_ensure(c, s, n_pairs)) != XFB_OK) return r;
  // copy-in stream
  const size_t dev_fs = (size_t)h * w;
  for (int b = 0; b < batch; ++b)
    XFB_CUDA_OK(c, cudaMemcpy2DAsync(s.d_gray + b * dev_fs, w, gray + b * frame_stride, stride, w, h, cudaMemcpyHostToDevice, c->s_h2d));
  XFB_CUDA_OK(c, cudaEventRecord(s.ev_h2d, c->s_h2d));
  // compute stream
  XFB_CUDA_OK(c, cudaStreamWaitEvent(c->stream, s.ev_h2d, 0));
  r = extract_device(c, s.d_gray, batch, dev_fs, h, w, w, topk, nms_thr, s.nvalid, s.xy, s.score, s.desc);
  if (r != XFB_OK) return r;
  int32_t* host_m[5] = {bi, bd, sd, ri, rd};
  if (n_pairs) {
    // the matches of this batch run on their own stream, so that the (latency-bound, low-occupancy) conv
    // kernels of the NEXT batch can share the SMs with the (1 CTA / SM) matcher kernels of this one
    XFB_CUDA_OK(c, cudaEventRecord(s.ev_ext, c->stream));
    XFB_CUDA_OK(c, cudaStreamWaitEvent(c->s_match, s.ev_ext, 0));
    XFB_CUDA_OK(c, cudaStreamWaitEvent(c->s_match, c->ev_match_free, 0));   // previous batch's matcher is done with the shared images
    int32_t* d[5];
    for (int i = 0; i < 5; ++i) d[i] = host_m[i] ? s.m[i] : nullptr;
    cudaStream_t keep = c->stream;
    c->stream = c->s_match;
    r = tc_match_frames(c, pairs, n_pairs, init_dist, d);
    c->stream = keep;
    if (r != XFB_OK) return r;
    XFB_CUDA_OK(c, cudaEventRecord(s.ev_comp, c->s_match));
    XFB_CUDA_OK(c, cudaEventRecord(c->ev_match_free, c->s_match));
  } else {
    XFB_CUDA_OK(c, cudaEventRecord(s.ev_comp, c->stream));
  }
  // copy-out stream
  XFB_CUDA_OK(c, cudaStreamWaitEvent(c->s_d2h, s.ev_comp, 0));
  const size_t K = topk;
  XFB_CUDA_OK(c, cudaMemcpyAsync(n_valid, s.nvalid, (size_t)batch * 4, cudaMemcpyDeviceToHost, c->s_d2h));
  XFB_CUDA_OK(c, cudaMemcpyAsync(kpt_xy, s.xy, batch * K * 2 * 4, cudaMemcpyDeviceToHost, c->s_d2h));
  XFB_CUDA_OK(c, cudaMemcpyAsync(score, s.score, batch * K * 4, cudaMemcpyDeviceToHost, c->s_d2h));
  XFB_CUDA_OK(c, cudaMemcpyAsync(desc, s.desc, batch * K * 64 * 4, cudaMemcpyDeviceToHost, c->s_d2h));
  for (int i = 0; i < 5; ++i)
    if (n_pairs && host_m[i]) XFB_CUDA_OK(c, cudaMemcpyAsync(host_m[i], s.m[i], (size_t)n_pairs * K * 4, cudaMemcpyDeviceToHost, c->s_d2h));
  XFB_CUDA_OK(c, cudaEventRecord(s.ev_d2h, c->s_d2h));
  // the next compute that reuses this slot's device buffers waits in xfb_submit -> xfb_wait(slot); nothing else aliases them
  s.pending = true;
  return XFB_OK;
}

int xfb_debug_match_error(xfb_ctx* c, const float* A, int n1, const float* B, int n2, float* max_err) {
  if (!c || !A || !B || !max_err || n1 <= 0 || n2 <= 0) return XFB_ERR_ARG;
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  int r = ensure_match_scratch(c, n1, n2, true);
  if (r != XFB_OK) return r;
  r = tc_ensure_generic(c, n1, n2);
  if (r != XFB_OK) return r;
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_a, A, (size_t)n1 * 256, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->m_b, B, (size_t)n2 * 256, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, cudaMemsetAsync(c->tc_dbg, 0, 4, c->stream));
  r = tc_matrix_generic(c, c->m_a, n1, c->m_b, n2, c->m_matrix, c->tc_dbg);
  if (r != XFB_OK) return r;
  XFB_CUDA_OK(c, cudaMemcpyAsync(max_err, c->tc_dbg, 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return XFB_OK;
}

int xfb_profile_enable(xfb_ctx* c, int enable) {
  if (!c) return XFB_ERR_ARG;
  c->prof = enable != 0;
  return XFB_OK;
}

int xfb_profile_read(xfb_ctx* c, float* ms, int32_t* count) {
  if (!c || !ms || !count) return XFB_ERR_ARG;
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < c->prof_tag.size(); ++i) {
    float t = 0.f;
    XFB_CUDA_OK(c, cudaEventElapsedTime(&t, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]));
    const int tag = c->prof_tag[i];
    if (tag >= 0 && tag < XFB_PROF_TAGS) { ms[tag] += t; count[tag] += 1; }
  }
  for (cudaEvent_t e : c->prof_ev) c->prof_pool.push_back(e);
  c->prof_ev.clear();
  c->prof_tag.clear();
  return XFB_OK;
}

const char* xfb_profile_tag_name(int tag) {
  if (tag >= 0 && tag < L_NUM) return kLayers[tag].ref_name;
  if (tag >= L_NUM && tag < P_NUM) return kStageNames[tag - L_NUM];
  return "";
}

// ---- introspection ------------------------------------------------------------------------------------
static int find_layer(const char* name) {
  for (int L = 0; L < L_NUM; ++L)
    if (std::strcmp(kLayers[L].ref_name, name) == 0) return L;
  return -1;
}

long xfb_debug_read(xfb_ctx* c, const char* name, int frame, float* host_out, size_t capacity, int32_t* dims) {
  if (!c || !name || !host_out || frame < 0 || frame >= c->B) return XFB_ERR_ARG;
  const float* src = nullptr;
  int h = 0, w = 0, ch = 1;
  const std::string nm(name);
  if (nm == "xn") { src = c->xn; h = c->H; w = c->W; }
  else if (nm == "x_pre") { src = c->xraw; h = c->H; w = c->W; }
  else if (nm == "avg4") { src = c->avg4; h = c->H >> 2; w = c->W >> 2; }
  else if (nm == "K1h") { src = c->k1h; h = c->H; w = c->W; }
  else if (nm == "pyramid_sum") { src = c->pyr; h = c->H >> 3; w = c->W >> 3; ch = 64; }
  else if (nm == "feats") { src = c->act[L_F_2]; h = c->H >> 3; w = c->W >> 3; ch = 64; }
  else if (nm == "H1") { src = c->act[L_HM_2]; h = c->H >> 3; w = c->W >> 3; }
  else {
    const int L = find_layer(name);
    if (L < 0 || L >= L_NUM_BN) { c->err = "debug_read: unknown tensor " + nm; return XFB_ERR_ARG; }
    src = c->act[L]; h = c->H >> kLayers[L].lvl_out; w = c->W >> kLayers[L].lvl_out; ch = kLayers[L].cout;
  }
  const size_t elems = (size_t)h * w * ch;
  if (dims) { dims[0] = h; dims[1] = w; dims[2] = ch; dims[3] = 0; }
  if (elems > capacity) { c->err = "debug_read: capacity too small"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  XFB_CUDA_OK(c, cudaMemcpy(host_out, src + (size_t)frame * elems, elems * 4, cudaMemcpyDeviceToHost));
  return (long)elems;
}

long xfb_debug_read_stats(xfb_ctx* c, const char* name, int frame, float* host_out, size_t capacity) {
  if (!c || !name || !host_out || frame < 0 || frame >= c->B) return XFB_ERR_ARG;
  const float *m = nullptr, *r = nullptr;
  int ch = 1;
  if (std::strcmp(name, "xn") == 0) { m = c->in_mean; r = c->in_rstd; }
  else {
    const int L = find_layer(name);
    if (L < 0 || L >= L_NUM_BN) { c->err = std::string("debug_read_stats: unknown layer ") + name; return XFB_ERR_ARG; }
    m = c->bn[L].mean; r = c->bn[L].rstd; ch = kLayers[L].cout;
  }
  if ((size_t)2 * ch > capacity) return XFB_ERR_ARG;
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  XFB_CUDA_OK(c, cudaMemcpy(host_out, m + (size_t)frame * ch, (size_t)ch * 4, cudaMemcpyDeviceToHost));
  XFB_CUDA_OK(c, cudaMemcpy(host_out + ch, r + (size_t)frame * ch, (size_t)ch * 4, cudaMemcpyDeviceToHost));
  return 2L * ch;
}

int xfb_debug_post(xfb_ctx* c, int H, int W, const float* feats, const float* H1, const float* K1h, int topk, float nms_thr,
                   int32_t* n_valid, float* kpt_xy, float* score, float* desc) {
  if (!c || !feats || !H1 || !K1h || !n_valid || !kpt_xy || !score || !desc) return XFB_ERR_ARG;
  if (H % 32 || W % 32 || H < 32 || W < 32 || H > c->max_h || W > c->max_w || topk < 1 || topk > c->max_topk) { c->err = "debug_post: bad size"; return XFB_ERR_ARG; }
  XFB_CUDA_OK(c, cudaSetDevice(c->device));
  c->B = 1; c->H = H; c->W = W; c->in_h = H; c->in_w = W;
  const size_t cells = (size_t)(H / 8) * (W / 8);
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->act[L_F_2], feats, cells * 64 * 4, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->act[L_HM_2], H1, cells * 4, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(c->k1h, K1h, (size_t)H * W * 4, cudaMemcpyHostToDevice, c->stream));
  XFB_CUDA_OK(c, launch_post(c, topk, nms_thr, c->o_nvalid, c->o_xy, c->o_score, c->o_desc));
  const size_t K = topk;
  XFB_CUDA_OK(c, cudaMemcpyAsync(n_valid, c->o_nvalid, 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(kpt_xy, c->o_xy, K * 2 * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(score, c->o_score, K * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaMemcpyAsync(desc, c->o_desc, K * 64 * 4, cudaMemcpyDeviceToHost, c->stream));
  XFB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
  return XFB_OK;
}

int xfb_debug_candidates(xfb_ctx* c, int frame) {
  if (!c || frame < 0 || frame >= c->max_batch) return XFB_ERR_ARG;
  int v = 0;
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return XFB_ERR_CUDA;
  if (cudaMemcpy(&v, c->cand_count_last + frame, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return XFB_ERR_CUDA;
  return v;
}

long xfb_launch_count(const xfb_ctx* c) { return c ? c->launches : 0; }

int xfb_debug_force_simt(xfb_ctx* c, int enable) {
  if (!c) return XFB_ERR_ARG;
  c->force_simt = enable != 0;
  return XFB_OK;
}

}  // extern "C"
