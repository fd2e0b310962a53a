// conv.cu -- FP32 direct convolution for every BasicLayer of XFeat (reference: BasicLayerImpl,
// src/XFeat.cc:7-28 = Conv2d(bias=false) -> BatchNorm2d(affine=false, TRAINING mode) -> ReLU).
//
// One kernel template serves all 3x3 / 1x1, stride 1 / 2 layers.  The train-mode BatchNorm is
// split across the producer/consumer pair so no activation is touched twice:
//   * the producer writes the RAW conv output (NHWC) and reduces per-(frame, channel) sum / sum of
//     squares: registers -> warp shuffles -> shared memory -> one FP64 partial per CTA in global
//     memory; the last CTA of a frame (ticket counter) folds the partials in a fixed order and
//     publishes mean and 1/sqrt(var + 1e-5).  Deterministic: no floating-point atomics.
//   * the consumer applies (x - mean) * rstd and ReLU while staging its input tile in shared memory.
// Statistics are per frame, so a batch of B frames in one launch equals B batch-1 forwards of the
// reference (SURVEY.md finding 1).
#include <cuda_runtime.h>

#include "xfb_internal.h"

namespace xfb {

const LayerSpec kLayers[L_NUM] = {
    {"block1.0", 1, 4, 3, 1, 0, 0},         {"block1.1", 4, 8, 3, 2, 0, 1},        {"block1.2", 8, 8, 3, 1, 1, 1},
    {"block1.3", 8, 24, 3, 2, 1, 2},        {"block2.0", 24, 24, 3, 1, 2, 2},      {"block2.1", 24, 24, 3, 1, 2, 2},
    {"block3.0", 24, 64, 3, 2, 2, 3},       {"block3.1", 64, 64, 3, 1, 3, 3},      {"block3.2", 64, 64, 1, 1, 3, 3},
    {"block4.0", 64, 64, 3, 2, 3, 4},       {"block4.1", 64, 64, 3, 1, 4, 4},      {"block4.2", 64, 64, 3, 1, 4, 4},
    {"block5.0", 64, 128, 3, 2, 4, 5},      {"block5.1", 128, 128, 3, 1, 5, 5},    {"block5.2", 128, 128, 3, 1, 5, 5},
    {"block5.3", 128, 64, 1, 1, 5, 5},      {"block_fusion.0", 64, 64, 3, 1, 3, 3}, {"block_fusion.1", 64, 64, 3, 1, 3, 3},
    {"heatmap_head.0", 64, 64, 1, 1, 3, 3}, {"heatmap_head.1", 64, 64, 1, 1, 3, 3}, {"keypoint_head.0", 64, 64, 1, 1, 3, 3},
    {"keypoint_head.1", 64, 64, 1, 1, 3, 3}, {"keypoint_head.2", 64, 64, 1, 1, 3, 3},
    {"block_fusion.2", 64, 64, 1, 1, 3, 3}, {"heatmap_head.2", 64, 1, 1, 1, 3, 3}, {"keypoint_head.3", 64, 65, 1, 1, 3, 3},
    {"skip1.1", 1, 24, 1, 1, 2, 2},
};

enum InMode { IN_PLAIN = 0, IN_BN = 1, IN_BN_SKIP = 2, IN_UNFOLD = 3 };
enum OutMode { OUT_STATS = 0, OUT_BIAS = 1 };

// Tile configuration of one layer shape.
//   TH x TW  output pixels per CTA;  PXT consecutive x-pixels and COUT/NCG channels per thread;
//   lane = cg + NCG * pixel-group, so the NCG channel groups of a pixel group sit in one warp
//   (weight reads are 16-byte vectors shared by the lanes of a channel group, activation reads
//   are broadcast across channel groups).  CK input channels are staged per pass.
template <int CIN_, int COUT_, int KS_, int S_, int TH_, int TW_, int PXT_, int NCG_, int CK_>
struct Cfg {
  static constexpr int CIN = CIN_, COUT = COUT_, KS = KS_, S = S_, TH = TH_, TW = TW_, PXT = PXT_, NCG = NCG_, CK = CK_;
  static constexpr int COT = COUT / NCG;
  static constexpr int NPGX = TW / PXT;
  static constexpr int NPG = TH * NPGX;
  static constexpr int NT = NPG * NCG;
  static constexpr int TIH = (TH - 1) * S + KS;
  static constexpr int TIW = (TW - 1) * S + KS;
  static constexpr int PAD = KS / 2;
  // pixel stride in shared memory == 1 (mod 4) words: the pixel groups of a warp hit distinct banks
  static constexpr int CKP = (CK == 1) ? 1 : (CK + ((5 - (CK % 4)) % 4));
  static constexpr int NXIN = (PXT - 1) * S + KS;
  static constexpr int SIN_WORDS = ((TIH * TIW * CKP + 3) / 4) * 4;
  static constexpr int SW_WORDS = KS * KS * CK * COUT;
  static constexpr int NWARP = NT / 32;
  static constexpr size_t SMEM_BYTES = sizeof(float) * (SIN_WORDS + SW_WORDS);
  static_assert(COUT % NCG == 0, "COUT must split evenly over channel groups");
  static_assert(TW % PXT == 0, "tile width must be a multiple of the per-thread pixel run");
  static_assert(NT % 32 == 0 && 32 % NCG == 0, "whole warps, channel groups inside a warp");
  static_assert(CIN % CK == 0, "input channel passes");
  static_assert(COT % 4 == 0, "vector weight loads");
};

__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }

template <class C, int INMODE, int OUTMODE>
__global__ void __launch_bounds__(C::NT) conv_bn_kernel(const ConvArgs a) {
  constexpr int CIN = C::CIN, COUT = C::COUT, KS = C::KS, S = C::S, PXT = C::PXT, NCG = C::NCG, COT = C::COT, CK = C::CK,
                CKP = C::CKP, TIH = C::TIH, TIW = C::TIW, NT = C::NT, PAD = C::PAD;
  extern __shared__ __align__(16) float smem[];
  float* sIn = smem;
  float* sW = smem + C::SIN_WORDS;

  const int t = threadIdx.x;
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * C::TH, ox0 = blockIdx.x * C::TW;
  const int cg = t % NCG;
  const int pg = t / NCG;
  const int pr = pg / C::NPGX;            // output row inside the tile
  const int px0 = (pg % C::NPGX) * PXT;   // first output column inside the tile
  const int iy_org = oy0 * S - PAD, ix_org = ox0 * S - PAD;

  float acc[PXT][COT];
#pragma unroll
  for (int p = 0; p < PXT; ++p)
#pragma unroll
    for (int q = 0; q < COT; ++q) acc[p][q] = 0.f;

  const float* in_b = (INMODE == IN_UNFOLD) ? a.in + (size_t)b * (a.Hin * 8) * a.full_w
                                            : a.in + (size_t)b * a.Hin * a.Win * CIN;

  for (int c0 = 0; c0 < CIN; c0 += CK) {
    if (c0) __syncthreads();
    // ---- stage the input tile, applying the producer's BatchNorm + ReLU on the fly -------------
    for (int idx = t; idx < TIH * TIW * CK; idx += NT) {
      const int c = idx % CK;
      const int xy = idx / CK;
      const int tx = xy % TIW, ty = xy / TIW;
      const int iy = iy_org + ty, ix = ix_org + tx;
      float v = 0.f;
      if (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) {
        const int ch = c0 + c;
        if (INMODE == IN_UNFOLD) {
          // XFeatModel::unfold2d(x, 8), src/XFeat.cc:124-133: channel = (y%8)*8 + x%8
          v = in_b[(size_t)(iy * 8 + (ch >> 3)) * a.full_w + ix * 8 + (ch & 7)];
        } else {
          v = in_b[((size_t)iy * a.Win + ix) * CIN + ch];
          if (INMODE == IN_BN || INMODE == IN_BN_SKIP) {
            v = fmaxf((v - a.in_mean[b * CIN + ch]) * a.in_rstd[b * CIN + ch], 0.f);
          }
          if (INMODE == IN_BN_SKIP) {
            // x1 + skip1(x), src/XFeat.cc:153; skip1 = AvgPool2d(4,4) + Conv2d(1,24,1) (:36-39)
            v += a.skip_avg[((size_t)b * a.Hin + iy) * a.Win + ix] * a.skip_w[ch] + a.skip_b[ch];
          }
        }
      }
      sIn[xy * CKP + c] = v;
    }
    // ---- stage the weight slice [KS*KS][CK][COUT] -----------------------------------------------
    for (int idx = t; idx < KS * KS * CK * COUT; idx += NT) {
      const int co = idx % COUT;
      const int kc = idx / COUT;
      const int c = kc % CK, k = kc / CK;
      sW[idx] = a.w[((size_t)k * CIN + c0 + c) * COUT + co];
    }
    __syncthreads();
    // ---- FMA core ---------------------------------------------------------------------------------
#pragma unroll 1
    for (int ky = 0; ky < KS; ++ky) {
      const float* row = sIn + ((pr * S + ky) * TIW + px0 * S) * CKP;
#pragma unroll 2
      for (int c = 0; c < CK; ++c) {
        float xin[C::NXIN];
#pragma unroll
        for (int i = 0; i < C::NXIN; ++i) xin[i] = row[i * CKP + c];
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const float4* wp = reinterpret_cast<const float4*>(sW + ((ky * KS + kx) * CK + c) * COUT + cg * COT);
          float wv[COT];
#pragma unroll
          for (int q = 0; q < COT / 4; ++q) {
            const float4 w4 = wp[q];
            wv[4 * q + 0] = w4.x; wv[4 * q + 1] = w4.y; wv[4 * q + 2] = w4.z; wv[4 * q + 3] = w4.w;
          }
#pragma unroll
          for (int p = 0; p < PXT; ++p)
#pragma unroll
            for (int q = 0; q < COT; ++q) acc[p][q] = fmaf(xin[p * S + kx], wv[q], acc[p][q]);
        }
      }
    }
  }

  // ---- epilogue -------------------------------------------------------------------------------------
  const int oy = oy0 + pr;
  float* out_b = a.out + (size_t)b * a.Hout * a.Wout * COUT;
  float s1[COT], s2[COT];
#pragma unroll
  for (int q = 0; q < COT; ++q) { s1[q] = 0.f; s2[q] = 0.f; }
#pragma unroll
  for (int p = 0; p < PXT; ++p) {
    const int ox = ox0 + px0 + p;
    if (oy < a.Hout && ox < a.Wout) {
      float4* dst = reinterpret_cast<float4*>(out_b + ((size_t)oy * a.Wout + ox) * COUT + cg * COT);
#pragma unroll
      for (int q = 0; q < COT / 4; ++q) {
        float4 v = make_float4(acc[p][4 * q], acc[p][4 * q + 1], acc[p][4 * q + 2], acc[p][4 * q + 3]);
        if (OUTMODE == OUT_BIAS) {
          const float4 bv = *reinterpret_cast<const float4*>(a.bias + cg * COT + 4 * q);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        }
        dst[q] = v;
      }
      if (OUTMODE == OUT_STATS) {
#pragma unroll
        for (int q = 0; q < COT; ++q) { s1[q] += acc[p][q]; s2[q] = fmaf(acc[p][q], acc[p][q], s2[q]); }
      }
    }
  }
  if (OUTMODE != OUT_STATS) return;

  // per-(frame, channel) statistics: lanes of one channel group -> warp -> CTA -> last CTA of the frame
  __syncthreads();  // everyone is done with sIn / sW; reuse shared memory below
  double* sRed = reinterpret_cast<double*>(smem);  // [NWARP][COUT][2]
#pragma unroll
  for (int q = 0; q < COT; ++q) {
#pragma unroll
    for (int off = NCG; off < 32; off <<= 1) {
      s1[q] += __shfl_xor_sync(0xffffffffu, s1[q], off);
      s2[q] += __shfl_xor_sync(0xffffffffu, s2[q], off);
    }
  }
  const int lane = t & 31, warp = t >> 5;
  if (lane < NCG) {
#pragma unroll
    for (int q = 0; q < COT; ++q) {
      sRed[(warp * COUT + lane * COT + q) * 2 + 0] = (double)s1[q];
      sRed[(warp * COUT + lane * COT + q) * 2 + 1] = (double)s2[q];
    }
  }
  __syncthreads();
  const int tiles = gridDim.x * gridDim.y;
  const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
  double* part_b = a.part + (size_t)b * tiles * COUT * 2;
  for (int c = t; c < COUT; c += NT) {
    double d1 = 0.0, d2 = 0.0;
    for (int w = 0; w < C::NWARP; ++w) { d1 += sRed[(w * COUT + c) * 2]; d2 += sRed[(w * COUT + c) * 2 + 1]; }
    part_b[((size_t)tile_id * COUT + c) * 2 + 0] = d1;
    part_b[((size_t)tile_id * COUT + c) * 2 + 1] = d2;
  }
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (t == 0) {
    const unsigned int prev = atomicAdd(a.ticket + b * XFB_TICKET_STRIDE, 1u);
    s_last = (prev == (unsigned int)(tiles - 1)) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA of frame b: fold all partials in a fixed order (slice-strided, then slice order)
  constexpr int NSL = (NT / COUT) > 0 ? (NT / COUT) : 1;   // slices per channel
  double* sFold = reinterpret_cast<double*>(smem);          // [NSL][COUT][2]
  for (int e = t; e < NSL * COUT; e += NT) {
    const int c = e % COUT, sl = e / COUT;
    double d1 = 0.0, d2 = 0.0;
    for (int i = sl; i < tiles; i += NSL) {
      d1 += __ldcg(part_b + ((size_t)i * COUT + c) * 2);
      d2 += __ldcg(part_b + ((size_t)i * COUT + c) * 2 + 1);
    }
    sFold[(sl * COUT + c) * 2] = d1;
    sFold[(sl * COUT + c) * 2 + 1] = d2;
  }
  __syncthreads();
  const double n = (double)a.Hout * (double)a.Wout;
  for (int c = t; c < COUT; c += NT) {
    double d1 = 0.0, d2 = 0.0;
    for (int sl = 0; sl < NSL; ++sl) { d1 += sFold[(sl * COUT + c) * 2]; d2 += sFold[(sl * COUT + c) * 2 + 1]; }
    const double mean = d1 / n;
    double var = d2 / n - mean * mean;   // biased variance, as BatchNorm uses for normalisation
    if (var < 0.0) var = 0.0;
    a.out_mean[b * COUT + c] = (float)mean;
    a.out_rstd[b * COUT + c] = (float)(1.0 / sqrt(var + 1e-5));
  }
  if (t == 0) a.ticket[b * XFB_TICKET_STRIDE] = 0u;  // re-arm for the next layer
}

// ---------------------------------------------------------------------------------------------------
// Layer shapes -> tile configurations.                   CIN COUT KS S  TH TW PXT NCG CK
using CfgB10 = Cfg<1, 4, 3, 1, 8, 64, 4, 1, 1>;        // 480x640   1->4
using CfgB11 = Cfg<4, 8, 3, 2, 16, 32, 4, 1, 4>;       // ->240x320 4->8
using CfgB12 = Cfg<8, 8, 3, 1, 16, 32, 4, 1, 8>;       // 240x320   8->8
using CfgB13 = Cfg<8, 24, 3, 2, 8, 32, 4, 2, 8>;       // ->120x160 8->24
using CfgB2x = Cfg<24, 24, 3, 1, 8, 32, 4, 2, 24>;     // 120x160   24->24
using CfgB30 = Cfg<24, 64, 3, 2, 4, 16, 4, 8, 24>;     // ->60x80   24->64
using CfgC33 = Cfg<64, 64, 3, 1, 8, 16, 8, 8, 16>;     // 60x80     64->64 3x3 (block3.1, fusion.0/.1)
using CfgC11 = Cfg<64, 64, 1, 1, 8, 16, 8, 8, 32>;     // 60x80     64->64 1x1
using CfgB40 = Cfg<64, 64, 3, 2, 4, 8, 2, 8, 16>;      // ->30x40   64->64
using CfgB4x = Cfg<64, 64, 3, 1, 4, 8, 2, 8, 16>;      // 30x40     64->64
using CfgB50 = Cfg<64, 128, 3, 2, 4, 4, 2, 16, 16>;    // ->15x20   64->128
using CfgB5x = Cfg<128, 128, 3, 1, 4, 4, 2, 16, 16>;   // 15x20     128->128
using CfgB53 = Cfg<128, 64, 1, 1, 4, 4, 2, 8, 32>;     // 15x20     128->64 1x1

template <class C, int INMODE, int OUTMODE>
static cudaError_t run(Ctx* c, const ConvArgs& a, int tag) {
  auto kern = conv_bn_kernel<C, INMODE, OUTMODE>;
  // the statistics fold reuses the staging buffers; make sure they are large enough
  constexpr size_t red = sizeof(double) * 2 * C::COUT * (C::NWARP > ((C::NT / C::COUT) > 0 ? (C::NT / C::COUT) : 1) ? C::NWARP : ((C::NT / C::COUT) > 0 ? (C::NT / C::COUT) : 1));
  constexpr size_t smem = C::SMEM_BYTES > red ? C::SMEM_BYTES : red;
  static unsigned long long attr_mask = 0;  // per device
  if (!((attr_mask >> c->device) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << c->device;
  }
  dim3 grid((a.Wout + C::TW - 1) / C::TW, (a.Hout + C::TH - 1) / C::TH, c->B);
  prof_begin(c, tag);
  kern<<<grid, C::NT, smem, c->stream>>>(a);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

template <class C>
static size_t part_need(int Hout, int Wout) {
  return (size_t)((Wout + C::TW - 1) / C::TW) * ((Hout + C::TH - 1) / C::TH) * C::COUT * 2;
}

size_t conv_part_elems(int H, int W) {
  size_t m = 0;
  auto upd = [&](size_t v) { if (v > m) m = v; };
  upd(part_need<CfgB10>(H, W));
  upd(part_need<CfgB11>(H / 2, W / 2));
  upd(part_need<CfgB12>(H / 2, W / 2));
  upd(part_need<CfgB13>(H / 4, W / 4));
  upd(part_need<CfgB2x>(H / 4, W / 4));
  upd(part_need<CfgB30>(H / 8, W / 8));
  upd(part_need<CfgC33>(H / 8, W / 8));
  upd(part_need<CfgC11>(H / 8, W / 8));
  upd(part_need<CfgB40>(H / 16, W / 16));
  upd(part_need<CfgB4x>(H / 16, W / 16));
  upd(part_need<CfgB50>(H / 32, W / 32));
  upd(part_need<CfgB5x>(H / 32, W / 32));
  upd(part_need<CfgB53>(H / 32, W / 32));
  return m;
}

cudaError_t launch_conv_layer(Ctx* c, int L) {
  const LayerSpec& sp = kLayers[L];
  ConvArgs a = {};
  a.Hin = c->H >> sp.lvl_in;  a.Win = c->W >> sp.lvl_in;
  a.Hout = c->H >> sp.lvl_out; a.Wout = c->W >> sp.lvl_out;
  a.w = c->w[L];
  a.bias = c->bias[L];
  a.out = c->act[L];
  a.part = c->part;
  a.ticket = c->ticket;
  a.full_w = c->W;
  if (L < L_NUM_BN) { a.out_mean = c->bn[L].mean; a.out_rstd = c->bn[L].rstd; }
  auto from = [&](int P) { a.in = c->act[P]; a.in_mean = c->bn[P].mean; a.in_rstd = c->bn[P].rstd; };
  switch (L) {
    case L_B1_0: a.in = c->xn; return run<CfgB10, IN_PLAIN, OUT_STATS>(c, a, L);
    case L_B1_1: from(L_B1_0); return run<CfgB11, IN_BN, OUT_STATS>(c, a, L);
    case L_B1_2: from(L_B1_1); return run<CfgB12, IN_BN, OUT_STATS>(c, a, L);
    case L_B1_3: from(L_B1_2); return run<CfgB13, IN_BN, OUT_STATS>(c, a, L);
    case L_B2_0:
      from(L_B1_3);
      a.skip_avg = c->avg4; a.skip_w = c->w[L_SKIP]; a.skip_b = c->bias[L_SKIP];
      return run<CfgB2x, IN_BN_SKIP, OUT_STATS>(c, a, L);
    case L_B2_1: from(L_B2_0); return run<CfgB2x, IN_BN, OUT_STATS>(c, a, L);
    case L_B3_0: from(L_B2_1); return run<CfgB30, IN_BN, OUT_STATS>(c, a, L);
    case L_B3_1: from(L_B3_0); return run<CfgC33, IN_BN, OUT_STATS>(c, a, L);
    case L_B3_2: from(L_B3_1); return run<CfgC11, IN_BN, OUT_STATS>(c, a, L);
    case L_B4_0: from(L_B3_2); return run<CfgB40, IN_BN, OUT_STATS>(c, a, L);
    case L_B4_1: from(L_B4_0); return run<CfgB4x, IN_BN, OUT_STATS>(c, a, L);
    case L_B4_2: from(L_B4_1); return run<CfgB4x, IN_BN, OUT_STATS>(c, a, L);
    case L_B5_0: from(L_B4_2); return run<CfgB50, IN_BN, OUT_STATS>(c, a, L);
    case L_B5_1: from(L_B5_0); return run<CfgB5x, IN_BN, OUT_STATS>(c, a, L);
    case L_B5_2: from(L_B5_1); return run<CfgB5x, IN_BN, OUT_STATS>(c, a, L);
    case L_B5_3: from(L_B5_2); return run<CfgB53, IN_BN, OUT_STATS>(c, a, L);
    case L_F_0: a.in = c->pyr; return run<CfgC33, IN_PLAIN, OUT_STATS>(c, a, L);
    case L_F_1: from(L_F_0); return run<CfgC33, IN_BN, OUT_STATS>(c, a, L);
    case L_F_2: from(L_F_1); return run<CfgC11, IN_BN, OUT_BIAS>(c, a, L);
    case L_HM_0: a.in = c->act[L_F_2]; return run<CfgC11, IN_PLAIN, OUT_STATS>(c, a, L);
    case L_HM_1: from(L_HM_0); return run<CfgC11, IN_BN, OUT_STATS>(c, a, L);
    case L_KP_0: a.in = c->xn; a.Hin = c->H >> 3; a.Win = c->W >> 3; return run<CfgC11, IN_UNFOLD, OUT_STATS>(c, a, L);
    case L_KP_1: from(L_KP_0); return run<CfgC11, IN_BN, OUT_STATS>(c, a, L);
    case L_KP_2: from(L_KP_1); return run<CfgC11, IN_BN, OUT_STATS>(c, a, L);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace xfb
