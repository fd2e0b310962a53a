// conv_small.cu -- block1 of XFeat (src/XFeat.cc:41-46: 1->4 s1, 4->8 s2, 8->8 s1, 8->24 s2, all 3x3 BasicLayers).
//
// These four layers hold 8 % of the MACs but the largest activations (4.9 MB / frame after block1.0): they are
// HBM-bound, and with <= 8 input channels there is no contraction worth a tensor core.  The kernel is written
// for bytes and instruction count, not FLOPs:
//   * lane = output column, each thread owns PY consecutive output rows x ALL output channels, so global stores
//     are full-line coalesced (COUT * 4 B per pixel, consecutive pixels contiguous in NHWC);
//   * the input tile is staged pixel-major [y][x][CIN] with 16-byte vector copies, the producer's train-mode
//     BatchNorm + ReLU applied in flight (SURVEY.md finding 1);
//   * per filter column the thread pulls its NYIN input pixels into registers once, and every weight vector
//     (broadcast LDS.128) is loaded once per (ky, kx, ci) and used for all PY rows.
// Statistics for the consumer's BatchNorm: registers -> shared memory -> FP64 partial per CTA -> last-CTA
// fixed-order fold (same deterministic protocol as conv.cu / conv_tc.cu).
#include <cuda_runtime.h>

#include "xfb_internal.h"

namespace xfb {

template <int CIN_, int COUT_, int S_, int PY_, int NW_>
struct SCfg {
  static constexpr int CIN = CIN_, COUT = COUT_, S = S_, PY = PY_, NW = NW_;
  static constexpr int NT = NW * 32;
  static constexpr int TW = 32, TH = NW * PY;                  // output tile
  static constexpr int TIW = (TW - 1) * S + 3, TIH = (TH - 1) * S + 3;
  static constexpr int NYIN = (PY - 1) * S + 3;                // input rows a thread touches
  static constexpr int Q = CIN >= 4 ? CIN / 4 : 1;             // 16-byte chunks per pixel
  static constexpr int IN_FLOATS = TIH * TIW * CIN;
  static constexpr int W_FLOATS = 9 * CIN * COUT;
  static constexpr int RED_FLOATS = NT * 2 * COUT;             // statistics scratch (reuses the tile)
  static constexpr size_t SMEM_BYTES = sizeof(float) * ((IN_FLOATS + 3) / 4 * 4 + W_FLOATS) > sizeof(float) * RED_FLOATS
                                           ? sizeof(float) * ((IN_FLOATS + 3) / 4 * 4 + W_FLOATS)
                                           : sizeof(float) * RED_FLOATS;
  static_assert(COUT % 4 == 0 && (CIN == 1 || CIN % 4 == 0), "channel vectors");
};

// STORE = false: statistics only (block1.0 when its output is never materialised).  FUSE0 = true (block1.1): the input tile is
// block1.0's output RECOMPUTED from the normalised image xn (1 -> 4 channels, 3x3; a.w0 = its weights, a.in = xn), with
// block1.0's BatchNorm + ReLU applied in flight -- the 4.9 MB / frame intermediate is neither written nor read.
template <class C, bool IN_BN, bool DIRECT, bool STORE = true, bool FUSE0 = false>
__global__ void __launch_bounds__(C::NT) conv_small_kernel(const ConvArgs a) {
  constexpr int CIN = C::CIN, COUT = C::COUT, S = C::S, PY = C::PY, NT = C::NT, TIW = C::TIW, TIH = C::TIH, NYIN = C::NYIN, Q = C::Q;
  extern __shared__ __align__(16) float smem[];
  float* sIn = smem;                                          // [TIH][TIW][CIN]
  float* sW = DIRECT ? smem + C::RED_FLOATS : smem + (C::IN_FLOATS + 3) / 4 * 4;   // [9][CIN][COUT]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * C::TH, ox0 = blockIdx.x * C::TW;
  const int iy_org = oy0 * S - 1, ix_org = ox0 * S - 1;
  const float* in_b = a.in + (size_t)b * a.Hin * a.Win * (FUSE0 ? 1 : CIN);

  // ---- stage weights and the input tile -----------------------------------------------------------------
  for (int i = t; i < C::W_FLOATS / 4; i += NT) reinterpret_cast<float4*>(sW)[i] = reinterpret_cast<const float4*>(a.w)[i];
  if (FUSE0) {
    // xn tile with one more ring of halo, then block1.0 on the fly: y0 = conv3x3(xn) (zero padding), relu(bn(y0)) -> sIn[ty][tx][0..3];
    // positions outside block1.0's output map are block1.1's zero padding
    float* sX = sW + C::W_FLOATS;                                   // [TIH + 2][TIW + 2]
    float* sW0 = sX + ((TIH + 2) * (TIW + 2) + 3) / 4 * 4;          // [9][4] (16-byte aligned)
    if (t < 36) sW0[t] = a.w0[t];
    for (int i = t; i < (TIH + 2) * (TIW + 2); i += NT) {
      const int ty = i / (TIW + 2), tx = i - ty * (TIW + 2);
      const int iy = iy_org - 1 + ty, ix = ix_org - 1 + tx;
      sX[i] = (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) ? in_b[(size_t)iy * a.Win + ix] : 0.f;
    }
    __syncthreads();
    const float4 m = *reinterpret_cast<const float4*>(a.in_mean + b * 4), r = *reinterpret_cast<const float4*>(a.in_rstd + b * 4);
    for (int i = t; i < TIH * TIW; i += NT) {
      const int ty = i / TIW, tx = i - ty * TIW;
      const int iy = iy_org + ty, ix = ix_org + tx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) {
        // same accumulation order as the statistics pass (kx outer, ky inner; fmaf chain from 0)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const float x = sX[(ty + ky) * (TIW + 2) + tx + kx];
            const float4 w4 = *reinterpret_cast<const float4*>(sW0 + (ky * 3 + kx) * 4);
            v.x = fmaf(x, w4.x, v.x); v.y = fmaf(x, w4.y, v.y); v.z = fmaf(x, w4.z, v.z); v.w = fmaf(x, w4.w, v.w);
          }
        v.x = fmaxf((v.x - m.x) * r.x, 0.f); v.y = fmaxf((v.y - m.y) * r.y, 0.f);
        v.z = fmaxf((v.z - m.z) * r.z, 0.f); v.w = fmaxf((v.w - m.w) * r.w, 0.f);
      }
      reinterpret_cast<float4*>(sIn)[i] = v;
    }
  } else if (DIRECT) {
    // no tile: every thread pulls its input pixels through L1 (neighbouring lanes / rows share the lines)
  } else if (CIN == 1) {
    for (int i = t; i < TIH * TIW; i += NT) {
      const int ty = i / TIW, tx = i - ty * TIW;
      const int iy = iy_org + ty, ix = ix_org + tx;
      float v = 0.f;
      if (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) {
        v = in_b[(size_t)iy * a.Win + ix];
        if (IN_BN) v = fmaxf((v - a.in_mean[b]) * a.in_rstd[b], 0.f);
      }
      sIn[i] = v;
    }
  } else {
    // NT is a multiple of Q, so a thread always handles the same channel chunk: BN parameters live in registers
    const int q = t % Q;
    float4 m = make_float4(0.f, 0.f, 0.f, 0.f), r = make_float4(1.f, 1.f, 1.f, 1.f);
    if (IN_BN) {
      m = *reinterpret_cast<const float4*>(a.in_mean + b * CIN + q * 4);
      r = *reinterpret_cast<const float4*>(a.in_rstd + b * CIN + q * 4);
    }
    for (int i = t; i < TIH * TIW * Q; i += NT) {
      const int pix = i / Q;
      const int ty = pix / TIW, tx = pix - ty * TIW;
      const int iy = iy_org + ty, ix = ix_org + tx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) {
        v = *reinterpret_cast<const float4*>(in_b + ((size_t)iy * a.Win + ix) * CIN + q * 4);
        if (IN_BN) {
          v.x = fmaxf((v.x - m.x) * r.x, 0.f); v.y = fmaxf((v.y - m.y) * r.y, 0.f);
          v.z = fmaxf((v.z - m.z) * r.z, 0.f); v.w = fmaxf((v.w - m.w) * r.w, 0.f);
        }
      }
      reinterpret_cast<float4*>(sIn)[i] = v;
    }
  }
  __syncthreads();

  // ---- compute: PY output rows x COUT channels per thread -----------------------------------------------
  // accumulators as channel PAIRS: FFMA2 (fma.rn.f32x2, sm_100) does two of the fp32 FMAs per issue slot -- same rounding, same order
  // per channel as the scalar form, half the instructions of the issue-bound inner loop
  float2 acc2[PY][COUT / 2];
#pragma unroll
  for (int p = 0; p < PY; ++p)
#pragma unroll
    for (int c = 0; c < COUT / 2; ++c) acc2[p][c] = make_float2(0.f, 0.f);
  const int row0 = warp * PY * S;                             // first input row of this thread inside the tile
  float4 bm[Q], br[Q];                                        // BN parameters of the producer (DIRECT path)
  if (DIRECT && IN_BN && CIN > 1) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      bm[q] = *reinterpret_cast<const float4*>(a.in_mean + b * CIN + q * 4);
      br[q] = *reinterpret_cast<const float4*>(a.in_rstd + b * CIN + q * 4);
    }
  }
#pragma unroll
  for (int kx = 0; kx < 3; ++kx) {
    float xin[NYIN][CIN];
    const float* col = sIn + ((size_t)row0 * TIW + lane * S + kx) * CIN;
    const int gx = ix_org + lane * S + kx;
#pragma unroll
    for (int iy = 0; iy < NYIN; ++iy) {
      if (DIRECT) {
        const int gy = iy_org + row0 + iy;
        const bool inb = gy >= 0 && gy < a.Hin && gx >= 0 && gx < a.Win;
        if (CIN == 1) {
          float v = 0.f;
          if (inb) {
            v = __ldg(in_b + (size_t)gy * a.Win + gx);
            if (IN_BN) v = fmaxf((v - a.in_mean[b]) * a.in_rstd[b], 0.f);
          }
          xin[iy][0] = v;
        } else {
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inb) {
              v = __ldg(reinterpret_cast<const float4*>(in_b + ((size_t)gy * a.Win + gx) * CIN + q * 4));
              if (IN_BN) {
                v.x = fmaxf((v.x - bm[q].x) * br[q].x, 0.f); v.y = fmaxf((v.y - bm[q].y) * br[q].y, 0.f);
                v.z = fmaxf((v.z - bm[q].z) * br[q].z, 0.f); v.w = fmaxf((v.w - bm[q].w) * br[q].w, 0.f);
              }
            }
            xin[iy][4 * q + 0] = v.x; xin[iy][4 * q + 1] = v.y; xin[iy][4 * q + 2] = v.z; xin[iy][4 * q + 3] = v.w;
          }
        }
      } else if (CIN == 1) xin[iy][0] = col[iy * TIW];
      else {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(col + (size_t)iy * TIW * CIN + q * 4);
          xin[iy][4 * q + 0] = v.x; xin[iy][4 * q + 1] = v.y; xin[iy][4 * q + 2] = v.z; xin[iy][4 * q + 3] = v.w;
        }
      }
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        float2 wv[COUT / 2];
        const float4* wp = reinterpret_cast<const float4*>(sW + ((ky * 3 + kx) * CIN + ci) * COUT);
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) {
          const float4 w4 = wp[c4];
          wv[2 * c4] = make_float2(w4.x, w4.y); wv[2 * c4 + 1] = make_float2(w4.z, w4.w);
        }
#pragma unroll
        for (int p = 0; p < PY; ++p) {
          const float x = xin[p * S + ky][ci];
          const float2 xx = make_float2(x, x);
#pragma unroll
          for (int c = 0; c < COUT / 2; ++c) {
            if (COUT >= 8) acc2[p][c] = __ffma2_rn(xx, wv[c], acc2[p][c]);
            else { acc2[p][c].x = fmaf(x, wv[c].x, acc2[p][c].x); acc2[p][c].y = fmaf(x, wv[c].y, acc2[p][c].y); }   // (block1.0: measured slower packed)
          }
        }
      }
    }
  }

  // ---- epilogue: coalesced store + per-channel sums ---------------------------------------------------------
  float acc[PY][COUT];
#pragma unroll
  for (int p = 0; p < PY; ++p)
#pragma unroll
    for (int c = 0; c < COUT / 2; ++c) { acc[p][2 * c] = acc2[p][c].x; acc[p][2 * c + 1] = acc2[p][c].y; }
  float* out_b = a.out + (size_t)b * a.Hout * a.Wout * COUT;
  const int ox = ox0 + lane;
  float s1[COUT], s2[COUT];
#pragma unroll
  for (int c = 0; c < COUT; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
#pragma unroll
  for (int p = 0; p < PY; ++p) {
    const int oy = oy0 + warp * PY + p;
    if (oy < a.Hout && ox < a.Wout) {
      if (STORE) {
        float4* dst = reinterpret_cast<float4*>(out_b + ((size_t)oy * a.Wout + ox) * COUT);
#pragma unroll
        for (int c4 = 0; c4 < COUT / 4; ++c4) dst[c4] = make_float4(acc[p][4 * c4], acc[p][4 * c4 + 1], acc[p][4 * c4 + 2], acc[p][4 * c4 + 3]);
      }
#pragma unroll
      for (int c = 0; c < COUT; ++c) { s1[c] += acc[p][c]; s2[c] = fmaf(acc[p][c], acc[p][c], s2[c]); }
    }
  }
  __syncthreads();                                            // the tile is dead: reuse it for the reduction
  float* sRed = smem;                                         // [NT][2 * COUT]
#pragma unroll
  for (int c = 0; c < COUT; ++c) { sRed[t * 2 * COUT + c] = s1[c]; sRed[t * 2 * COUT + COUT + c] = s2[c]; }
  __syncthreads();
  // column sums in a fixed order: thread j < 2*COUT*NSEG sums a segment of rows of column j % (2*COUT)
  constexpr int NV = 2 * COUT;
  constexpr int NSEG = NT / NV > 0 ? NT / NV : 1;
  constexpr int ROWS = NT / NSEG;
  __shared__ double sSeg[256 * 2];                            // [NSEG][NV] (NSEG * NV <= NT <= 256)
  if (t < NSEG * NV) {
    const int v = t % NV, sg = t / NV;
    double d = 0.0;
    const int r1 = (sg == NSEG - 1) ? NT : (sg + 1) * ROWS;
    for (int rr = sg * ROWS; rr < r1; ++rr) d += (double)sRed[rr * NV + v];
    sSeg[sg * NV + v] = d;
  }
  __syncthreads();
  const int tiles = gridDim.x * gridDim.y;
  const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
  double* part_b = a.part + (size_t)b * tiles * COUT * 2;
  if (t < NV) {
    double d = 0.0;
    for (int sg = 0; sg < NSEG; ++sg) d += sSeg[sg * NV + t];
    const int c = t % COUT, which = t / COUT;                 // 0 = sum, 1 = sum of squares
    part_b[((size_t)tile_id * COUT + c) * 2 + which] = d;
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
  // last CTA of the frame: fixed-order fold (slice-strided, then slice order)
  constexpr int NSL = NT / COUT > 0 ? NT / COUT : 1;
  double* sFold = reinterpret_cast<double*>(smem);            // [NSL][COUT][2]
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
    double var = d2 / n - mean * mean;
    if (var < 0.0) var = 0.0;
    a.out_mean[b * COUT + c] = (float)mean;
    a.out_rstd[b * COUT + c] = (float)(1.0 / sqrt(var + 1e-5));
  }
  if (t == 0) a.ticket[b * XFB_TICKET_STRIDE] = 0u;
}

#ifndef XFB_B11_NW
#define XFB_B11_NW 8
#endif
#ifndef XFB_B11_DIRECT
#define XFB_B11_DIRECT true
#endif
#ifndef XFB_B13_DIRECT
#define XFB_B13_DIRECT true
#endif
#ifndef XFB_B12_DIRECT
#define XFB_B12_DIRECT false
#endif
#ifndef XFB_B12_NW
#define XFB_B12_NW 4
#endif
#ifndef XFB_B13_NW
#define XFB_B13_NW 4
#endif
//                    CIN COUT S PY NW
using SB10 = SCfg<1, 4, 1, 4, 8>;     // block1.0  480x640            tile 32 x 32
using SB11 = SCfg<4, 8, 2, 4, XFB_B11_NW>;     // block1.1  -> 240x320         tile 32 x 32 (input 65 x 65 x 4)
using SB12 = SCfg<8, 8, 1, 4, XFB_B12_NW>;     // block1.2  240x320            tile 32 x 4*NW
using SB13 = SCfg<8, 24, 2, 2, XFB_B13_NW>;    // block1.3  -> 120x160         tile 32 x 2*NW

template <class C, bool IN_BN, bool DIRECT, bool STORE = true, bool FUSE0 = false>
static cudaError_t run_small(Ctx* c, const ConvArgs& a, int tag) {
  auto kern = conv_small_kernel<C, IN_BN, DIRECT, STORE, FUSE0>;
  // DIRECT needs no input tile: reduction scratch + weights only (more CTAs per SM); FUSE0 adds the xn tile and block1.0's weights
  const size_t smem = DIRECT ? sizeof(float) * (C::RED_FLOATS + C::W_FLOATS)
                             : (FUSE0 ? C::SMEM_BYTES + sizeof(float) * ((C::TIH + 2) * (C::TIW + 2) + 36 + 8) : C::SMEM_BYTES);
  static unsigned long long attr_mask = 0;
  if (!((attr_mask >> c->device) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 49152 ? smem : 49152));
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

size_t conv_small_part_elems(int H, int W) {
  size_t m = 0;
  auto upd = [&](int h, int w, int th, int cout) {
    const size_t v = (size_t)((w + 31) / 32) * ((h + th - 1) / th) * cout * 2;
    if (v > m) m = v;
  };
  upd(H, W, SB10::TH, 4); upd(H / 2, W / 2, SB11::TH, 8); upd(H / 2, W / 2, SB12::TH, 8); upd(H / 4, W / 4, SB13::TH, 24);
  return m;
}

bool conv_small_handles(int L) { return L == L_B1_0 || L == L_B1_1 || L == L_B1_2 || L == L_B1_3; }

cudaError_t launch_conv_small_layer(Ctx* c, int L) {
  const LayerSpec& sp = kLayers[L];
  ConvArgs a = {};
  a.Hin = c->H >> sp.lvl_in; a.Win = c->W >> sp.lvl_in;
  a.Hout = c->H >> sp.lvl_out; a.Wout = c->W >> sp.lvl_out;
  a.w = c->w[L];
  a.out = c->act[L];
  a.part = c->part; a.ticket = c->ticket;
  a.out_mean = c->bn[L].mean; a.out_rstd = c->bn[L].rstd;
  auto from = [&](int P) { a.in = c->act[P]; a.in_mean = c->bn[P].mean; a.in_rstd = c->bn[P].rstd; };
  switch (L) {
    // DIRECT (inputs through L1, no staged tile) measured faster except for the 8->8 stride-1 layer
    // (18 two-vector pixel loads per thread): 0.070 / 0.133 / 0.128 / 0.110 ms per 32-frame batch
    // XFB_B1_FUSE=1: block1.0 -> block1.1 fused by recomputation (block1.0 only produces its BatchNorm statistics, block1.1 rebuilds
    // block1.0's output tile from xn in shared memory).  Measured on the B200: 0.060 + 0.152 ms against 0.068 + 0.133 ms for the two
    // kernels -- the recomputation costs more than the 9.8 MB / frame of traffic it saves, so the two-kernel form stays the default.
    case L_B1_0: a.in = c->xn; return c->b1_fuse ? run_small<SB10, false, true, false>(c, a, L) : run_small<SB10, false, true>(c, a, L);
    case L_B1_1:
      if (c->b1_fuse) {
        a.in = c->xn; a.in_mean = c->bn[L_B1_0].mean; a.in_rstd = c->bn[L_B1_0].rstd; a.w0 = c->w[L_B1_0];
        a.Hin = c->H; a.Win = c->W;
        return run_small<SB11, true, false, true, true>(c, a, L);
      }
      from(L_B1_0);
      return run_small<SB11, true, XFB_B11_DIRECT>(c, a, L);
    case L_B1_2: from(L_B1_1); return run_small<SB12, true, XFB_B12_DIRECT>(c, a, L);
    case L_B1_3: from(L_B1_2); return run_small<SB13, true, XFB_B13_DIRECT>(c, a, L);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace xfb
