// conv_tc.cu -- the dense convolutions of XFeat (Cin >= 24) as implicit GEMMs on tcgen05 tensor cores.
//
// Same contract as conv.cu (reference: BasicLayerImpl, src/XFeat.cc:7-28 -- conv, TRAIN-mode BatchNorm,
// ReLU): the producer's BN+ReLU (+ skip1 add, + unfold2d) is applied while the input tile is staged, the raw
// conv output is written NHWC, and per-(frame, channel) sum / sum-of-squares go through the fixed-order FP64
// fold.  What changes is the contraction: D[128 pixels x COUT] += A[128 x CIN] * W[COUT x CIN]^T per filter
// tap on the tensor cores (UTCHMMA, fp32 accumulators in TMEM), as a 3xTF32 split (x = hi + lo exactly;
// hi*hi + hi*lo + lo*hi), which keeps fp32-level accuracy (parity target: descriptors within 1e-4 of the
// fp32 reference; measured 5e-6).
//
// No im2col.  An output tile is 8 wide x 16 tall = 128 pixels = the 128 TMEM lanes.  The halo tile is staged
// ONCE per channel phase in shared memory as channel-chunk planes [cin/4][pixel][4 floats].  That is exactly
// the canonical K-major no-swizzle UMMA operand layout with 16 B between the 8 pixels of a row group (= one
// output row), SBO = tile_width * 16 B between output rows and LBO = plane stride between 16-byte channel
// chunks -- so the A operand of filter tap (ky, kx) is the SAME buffer with the descriptor start address
// advanced by (ky * tile_width + kx) * 16 B.  Stride-2 layers stage four parity planes (even/odd rows x
// even/odd columns) so that consecutive output pixels are again 16 B apart.  Weights are pre-split and
// pre-tiled per (channel phase, tap) at xfb_create and streamed with cp.async.bulk through an mbarrier ring.
//
// Warp roles (320 threads): warps 0-7 stage the input tile (4 independent 16-byte loads in flight per
// thread), later run the epilogue (TMEM lane quadrant = warp id % 4, two warps per quadrant split the
// columns); warp 8 = weight loader; warp 9 = TMEM allocator + single-thread MMA issuer.
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "tc_ptx.cuh"
#include "xfb_internal.h"

namespace xfb {

enum TcInMode { TIN_PLAIN = 0, TIN_BN = 1, TIN_BN_SKIP = 2, TIN_UNFOLD = 3 };
enum TcOutMode { TOUT_STATS = 0, TOUT_BIAS = 1, TOUT_KPSOFTMAX = 2 };

template <int CIN_, int COUT_, int KS_, int S_, int CSTAGE_>
struct TcCfg {
  static constexpr int CIN = CIN_, COUT = COUT_, KS = KS_, S = S_, CSTAGE = CSTAGE_;
  static constexpr int NP = (COUT + 15) / 16 * 16 < 32 ? 32 : (COUT + 15) / 16 * 16;   // UMMA N (M = 128 needs N % 16 == 0)
  static constexpr int TW = 8, TH = 16, PAD = KS / 2;
  static constexpr int NSUB = S == 2 ? 4 : 1;              // parity planes
  static constexpr int WT = S == 2 ? TW + 1 : TW + 2 * PAD;
  static constexpr int HT = S == 2 ? TH + 1 : TH + 2 * PAD;
  static constexpr int NPIX = WT * HT;
  static constexpr int KCS = CSTAGE / 4;                   // 16-byte channel chunks per phase
  static constexpr int PS = NPIX * 4 + 4;                  // plane stride in floats (padded by one chunk: bank spread)
  static constexpr int SUB_FLOATS = KCS * PS;
  static constexpr int IN_FLOATS = NSUB * SUB_FLOATS;      // one of hi / lo
  static constexpr int TAPS = KS * KS;
  static constexpr int NPHASE = CIN / CSTAGE;
  static constexpr int UNITS = NPHASE * TAPS;              // weight stage units (phase, tap)
  static constexpr int W_UNIT_FLOATS = 2 * CSTAGE * NP;    // hi + lo image of one unit
  static constexpr int OUT_LD = COUT + 4;                  // epilogue staging row stride (floats)
  static constexpr uint32_t LBO_A = PS * 4, SBO_A = WT * 16;
  static constexpr uint32_t LBO_B = (NP / 8) * 128, SBO_B = 128;
  static constexpr uint32_t TMEM_COLS = NP <= 32 ? 32 : (NP <= 64 ? 64 : (NP <= 128 ? 128 : 256));
  static constexpr size_t SMEM_IN = sizeof(float) * 2 * IN_FLOATS;
  static constexpr size_t SMEM_OUT = sizeof(float) * 128 * OUT_LD;
  static constexpr size_t SMEM_MAIN = ((SMEM_IN > SMEM_OUT ? SMEM_IN : SMEM_OUT) + 127) / 128 * 128;   // staging tile, reused by the epilogue
  static constexpr size_t SMEM_TAIL = 512 * 8 + 256;       // statistics scratch + barriers
  static constexpr int NSTAGE_FIT = (int)((232448 - SMEM_MAIN - SMEM_TAIL) / (sizeof(float) * W_UNIT_FLOATS));
  static constexpr int NSTAGE = UNITS < 3 ? UNITS : (NSTAGE_FIT < 3 ? NSTAGE_FIT : 3);
  static constexpr size_t SMEM_BYTES = SMEM_MAIN + sizeof(float) * NSTAGE * W_UNIT_FLOATS + SMEM_TAIL;
  static constexpr int NSLICE = (256 / COUT >= 8) ? 8 : ((256 / COUT >= 4) ? 4 : ((256 / COUT >= 2) ? 2 : 1));
  static_assert(CSTAGE % 8 == 0 && CIN % CSTAGE == 0 && COUT % 4 == 0 && NP <= 256, "shape");
  static_assert(S == 1 || KS == 3, "stride 2 is implemented for 3x3 only");
  static_assert(NSTAGE >= 1 && SMEM_BYTES <= 232448, "shared memory budget");
};

struct ConvTcArgs {
  const float* in;        // NHWC raw producer output (or plain values), or xn for TIN_UNFOLD
  const float* wimg;      // [phase][tap][hi | lo][cstage/4][np/8][8][4]
  const float* bias;
  float* out;
  const float* in_mean; const float* in_rstd;
  const float* skip_avg; const float* skip_w; const float* skip_b;
  double* part; unsigned int* ticket; float* out_mean; float* out_rstd;
  int Hin, Win, Hout, Wout;
  int full_w;             // TIN_UNFOLD: width of xn
};

constexpr int CTC_PROD = 256;               // producer / epilogue threads (8 warps)
constexpr int CTC_THREADS = CTC_PROD + 64;  // + weight-loader warp + MMA warp
constexpr int CTC_UNR = 4;                  // independent global loads in flight per producer thread

template <class C, int INMODE, int OUTMODE>
__global__ void __launch_bounds__(CTC_THREADS, 1) conv_tc_kernel(const ConvTcArgs a) {
  constexpr int CIN = C::CIN, COUT = C::COUT, KCS = C::KCS, PS = C::PS, WT = C::WT, NPIX = C::NPIX, TAPS = C::TAPS, NSTAGE = C::NSTAGE,
                NPHASE = C::NPHASE, NSUB = C::NSUB;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* sInHi = reinterpret_cast<float*>(smem_raw);
  float* sInLo = sInHi + C::IN_FLOATS;
  float* sOut = reinterpret_cast<float*>(smem_raw);                                    // reused after the MMAs are done
  float* sW = reinterpret_cast<float*>(smem_raw + C::SMEM_MAIN);
  double* sStat = reinterpret_cast<double*>(sW + NSTAGE * C::W_UNIT_FLOATS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 512);
  uint64_t* bar_in = bars + 0;            // input tile staged (128 producer arrivals per phase)
  uint64_t* bar_free = bars + 1;          // tensor core finished reading the staged tile (per phase)
  uint64_t* bar_wfull = bars + 2;         // [NSTAGE]
  uint64_t* bar_wempty = bars + 5;        // [NSTAGE]
  uint64_t* bar_acc = bars + 8;           // accumulator complete
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 9);
  __shared__ unsigned int s_last;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * C::TH, ox0 = blockIdx.x * C::TW;

  if (t == 0) {
    mbar_init(bar_in, CTC_PROD);
    mbar_init(bar_free, 1);
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bar_wfull + s, 1); mbar_init(bar_wempty + s, 1); }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc(s_tmem, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < 8) {
    // ===== stage the halo tile: BN + ReLU of the producer, exact hi/lo split, channel-chunk planes =====
    const float* in_b = (INMODE == TIN_UNFOLD) ? a.in + (size_t)b * (a.Hin * 8) * a.full_w : a.in + (size_t)b * a.Hin * a.Win * CIN;
    constexpr int ITEMS = NSUB * NPIX * KCS;
    for (int ph = 0; ph < NPHASE; ++ph) {
      if (ph > 0) mbar_wait(bar_free, (ph - 1) & 1);      // the previous phase's MMAs are done with the buffer
      for (int base = 0; base < ITEMS; base += CTC_PROD * CTC_UNR) {
        float4 v[CTC_UNR];
        float av[CTC_UNR];
        int off[CTC_UNR], chn[CTC_UNR];
        // issue all loads of this batch first (memory-level parallelism), then transform and store
#pragma unroll
        for (int u = 0; u < CTC_UNR; ++u) {
          const int idx = base + u * CTC_PROD + t;
          v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          av[u] = 0.f;
          off[u] = -1;
          chn[u] = -1;
          if (idx < ITEMS) {
            const int kc = idx % KCS;
            const int rest = idx / KCS;
            const int pix = rest % NPIX, sub = rest / NPIX;
            int iy, ix;
            if (C::S == 1) { iy = oy0 - C::PAD + pix / WT; ix = ox0 - C::PAD + pix % WT; }
            else { iy = 2 * (oy0 - 1 + pix / WT) + (sub >> 1); ix = 2 * (ox0 - 1 + pix % WT) + (sub & 1); }
            const int ch = ph * C::CSTAGE + kc * 4;
            off[u] = sub * C::SUB_FLOATS + kc * PS + pix * 4;
            if (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) {
              chn[u] = ch;
              if (INMODE == TIN_UNFOLD) {
                // XFeatModel::unfold2d(x, 8), src/XFeat.cc:124-133: channel c = (y % 8) * 8 + x % 8
                v[u] = *reinterpret_cast<const float4*>(in_b + (size_t)(iy * 8 + (ch >> 3)) * a.full_w + ix * 8 + (ch & 7));
              } else {
                v[u] = *reinterpret_cast<const float4*>(in_b + ((size_t)iy * a.Win + ix) * CIN + ch);
                if (INMODE == TIN_BN_SKIP) av[u] = a.skip_avg[((size_t)b * a.Hin + iy) * a.Win + ix];
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < CTC_UNR; ++u) {
          if (off[u] < 0) continue;
          float4 x = v[u];
          if (chn[u] >= 0 && (INMODE == TIN_BN || INMODE == TIN_BN_SKIP)) {
            const float4 m = *reinterpret_cast<const float4*>(a.in_mean + b * CIN + chn[u]);
            const float4 r = *reinterpret_cast<const float4*>(a.in_rstd + b * CIN + chn[u]);
            x.x = fmaxf((x.x - m.x) * r.x, 0.f); x.y = fmaxf((x.y - m.y) * r.y, 0.f);
            x.z = fmaxf((x.z - m.z) * r.z, 0.f); x.w = fmaxf((x.w - m.w) * r.w, 0.f);
            if (INMODE == TIN_BN_SKIP) {
              // x1 + skip1(x), src/XFeat.cc:153; skip1 = AvgPool2d(4,4) + Conv2d(1,24,1) (:36-39)
              const float4 sw = *reinterpret_cast<const float4*>(a.skip_w + chn[u]);
              const float4 sb = *reinterpret_cast<const float4*>(a.skip_b + chn[u]);
              x.x += av[u] * sw.x + sb.x; x.y += av[u] * sw.y + sb.y; x.z += av[u] * sw.z + sb.z; x.w += av[u] * sw.w + sb.w;
            }
          }
          float4 hi, lo;
          tf32_split(x.x, hi.x, lo.x); tf32_split(x.y, hi.y, lo.y); tf32_split(x.z, hi.z, lo.z); tf32_split(x.w, hi.w, lo.w);
          *reinterpret_cast<float4*>(sInHi + off[u]) = hi;
          *reinterpret_cast<float4*>(sInLo + off[u]) = lo;
        }
      }
      fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
      mbar_arrive(bar_in);
    }
  } else if (warp == 8) {
    // ===== weight loader (one elected lane): one bulk copy per (phase, tap) unit through the stage ring =====
    if (elect_one_sync()) {
      for (int u = 0; u < C::UNITS; ++u) {
        const int s = u % NSTAGE;
        if (u >= NSTAGE) mbar_wait(bar_wempty + s, ((u / NSTAGE) - 1) & 1);
        mbar_expect_tx(bar_wfull + s, C::W_UNIT_FLOATS * 4);
        bulk_g2s(sW + (size_t)s * C::W_UNIT_FLOATS, a.wimg + (size_t)u * C::W_UNIT_FLOATS, C::W_UNIT_FLOATS * 4, bar_wfull + s);
      }
    }
  } else {
    // ===== MMA issuer: the whole warp runs the (warp-uniform) loop and ONE elected lane issues, so that every tcgen05
    // instruction compiles to a single UTCHMMA / UTCBAR; descriptors are a constant plus 16-byte-unit offsets =====
    {
      constexpr uint32_t IDESC = umma_idesc_tf32(128, C::NP);
      const uint64_t da_hi0 = umma_desc_kmajor(smem_u32(sInHi), C::LBO_A, C::SBO_A);
      const uint64_t da_lo0 = umma_desc_kmajor(smem_u32(sInLo), C::LBO_A, C::SBO_A);
      const uint64_t db0 = umma_desc_kmajor(smem_u32(sW), C::LBO_B, C::SBO_B);
      constexpr uint64_t KA = (2u * C::LBO_A) >> 4, KB = (2u * C::LBO_B) >> 4;      // one K = 8 step (two 16-byte chunks)
      constexpr uint64_t W_LO = ((uint64_t)C::CSTAGE * C::NP * 4) >> 4;              // lo image behind the hi image
      constexpr uint64_t W_STAGE = ((uint64_t)C::W_UNIT_FLOATS * 4) >> 4;
      const bool leader = elect_one_sync();
      int u = 0;
      for (int ph = 0; ph < NPHASE; ++ph) {
        mbar_wait(bar_in, ph & 1);
        tc_fence_after();
        for (int tap = 0; tap < TAPS; ++tap, ++u) {
          const int s = u % NSTAGE;
          mbar_wait(bar_wfull + s, (u / NSTAGE) & 1);
          tc_fence_after();
          // shifted window of the staged tile (stride 2: pick the parity plane of this tap)
          const int ky = tap / C::KS, kx = tap % C::KS;
          uint32_t tap_off;
          if (C::S == 1) tap_off = (uint32_t)(ky * WT + kx) * 16u;
          else {
            const int py = (ky == 1) ? 0 : 1, dy = (ky == 0) ? 0 : 1, px = (kx == 1) ? 0 : 1, dx = (kx == 0) ? 0 : 1;
            tap_off = (uint32_t)((py * 2 + px) * C::SUB_FLOATS) * 4u + (uint32_t)(dy * WT + dx) * 16u;
          }
          if (leader) {
            const uint64_t dah0 = da_hi0 + (tap_off >> 4), dal0 = da_lo0 + (tap_off >> 4);
            const uint64_t dbh0 = db0 + (uint64_t)s * W_STAGE, dbl0 = dbh0 + W_LO;
#pragma unroll
            for (int k8 = 0; k8 < C::CSTAGE / 8; ++k8) {
              umma_tf32(tmem_base, dah0 + k8 * KA, dbh0 + k8 * KB, IDESC, (u > 0 || k8 > 0) ? 1u : 0u);
              umma_tf32(tmem_base, dah0 + k8 * KA, dbl0 + k8 * KB, IDESC, 1u);
              umma_tf32(tmem_base, dal0 + k8 * KA, dbh0 + k8 * KB, IDESC, 1u);
            }
            umma_commit(bar_wempty + s);
          }
          __syncwarp();
        }
        if (leader) umma_commit(bar_free);            // staged tile may be overwritten by the next channel phase
        __syncwarp();
      }
      if (leader) umma_commit(bar_acc);
      __syncwarp();
    }
  }

  if (OUTMODE == TOUT_KPSOFTMAX) {
    // ===== keypoint_head.3 epilogue: + bias, softmax over the 65 logits, drop the dustbin, 8x8 fold =====
    // (src/XFeat.cc:85-90, XFextractor::getKptsHeatmap src/XFextractor.cc:204-217); one thread per cell
    if (warp < 4) {
      mbar_wait(bar_acc, 0);
      tc_fence_after();
      const int p = warp * 32 + lane;
      float v[96];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 0u, v);
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32u, v + 32);
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + 64u, v + 64);
      float mx = -3.4e38f;
#pragma unroll
      for (int c = 0; c < 65; ++c) { v[c] += a.bias[c]; mx = fmaxf(mx, v[c]); }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 65; ++c) { v[c] = expf(v[c] - mx); sum += v[c]; }
      const int Y = oy0 + (p >> 3), X = ox0 + (p & 7);
      if (Y < a.Hout && X < a.Wout) {
        float* dst = a.out + (size_t)b * (a.Hout * 8) * a.full_w + (size_t)(Y * 8) * a.full_w + X * 8;
#pragma unroll
        for (int ry = 0; ry < 8; ++ry) {
          const float4 lo4 = make_float4(v[ry * 8 + 0] / sum, v[ry * 8 + 1] / sum, v[ry * 8 + 2] / sum, v[ry * 8 + 3] / sum);
          const float4 hi4 = make_float4(v[ry * 8 + 4] / sum, v[ry * 8 + 5] / sum, v[ry * 8 + 6] / sum, v[ry * 8 + 7] / sum);
          *reinterpret_cast<float4*>(dst + (size_t)ry * a.full_w) = lo4;
          *reinterpret_cast<float4*>(dst + (size_t)ry * a.full_w + 4) = hi4;
        }
      }
      tc_fence_before();
    }
  } else if (warp < 8) {
    // ===== epilogue: TMEM -> registers -> smem staging -> coalesced NHWC store + channel statistics =====
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int quad = warp & 3, half = warp >> 2;     // two warps per TMEM lane quadrant, each takes half of the columns
    const int p = quad * 32 + lane;                   // pixel of the tile = TMEM lane
    constexpr int NCH = C::NP / 32;                   // 32-column chunks
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
      if ((NCH == 1 && half == 0) || (NCH > 1 && (ci * 2) / NCH == half)) {
        const int c0 = ci * 32;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          if (c0 + q < COUT) {
            float4 o = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            if (OUTMODE == TOUT_BIAS) {
              const float4 bv = *reinterpret_cast<const float4*>(a.bias + c0 + q);
              o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
            }
            *reinterpret_cast<float4*>(sOut + p * C::OUT_LD + c0 + q) = o;
          }
        }
      }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // coalesced store: the 8 pixels of an output row are 8 * COUT contiguous floats in NHWC
    float* out_b = a.out + (size_t)b * a.Hout * a.Wout * COUT;
    constexpr int V4_PER_PIX = COUT / 4;
    for (int idx = t; idx < 128 * V4_PER_PIX; idx += CTC_PROD) {
      const int pp = idx / V4_PER_PIX, c4 = idx % V4_PER_PIX;
      const int oy = oy0 + (pp >> 3), ox = ox0 + (pp & 7);
      if (oy < a.Hout && ox < a.Wout)
        *reinterpret_cast<float4*>(out_b + ((size_t)oy * a.Wout + ox) * COUT + c4 * 4) = *reinterpret_cast<const float4*>(sOut + pp * C::OUT_LD + c4 * 4);
    }
    if (OUTMODE == TOUT_STATS) {
      // per-channel sums over the valid pixels of the tile: COUT channels x NSLICE pixel slices
      constexpr int NSLICE = C::NSLICE, PIX_PER = 128 / NSLICE;
      for (int item = t; item < COUT * NSLICE; item += CTC_PROD) {
        const int c = item % COUT, sl = item / COUT;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
        for (int pp = sl * PIX_PER; pp < (sl + 1) * PIX_PER; ++pp) {
          const int oy = oy0 + (pp >> 3), ox = ox0 + (pp & 7);
          if (oy < a.Hout && ox < a.Wout) {
            const float x = sOut[pp * C::OUT_LD + c];
            s1 += x;
            s2 = fmaf(x, x, s2);
          }
        }
        sStat[(sl * COUT + c) * 2] = (double)s1;
        sStat[(sl * COUT + c) * 2 + 1] = (double)s2;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int tiles = gridDim.x * gridDim.y;
      const int tile_id = blockIdx.y * gridDim.x + blockIdx.x;
      double* part_b = a.part + (size_t)b * tiles * COUT * 2;
      for (int c = t; c < COUT; c += CTC_PROD) {
        double d1 = 0.0, d2 = 0.0;
        for (int sl = 0; sl < NSLICE; ++sl) { d1 += sStat[(sl * COUT + c) * 2]; d2 += sStat[(sl * COUT + c) * 2 + 1]; }
        part_b[((size_t)tile_id * COUT + c) * 2] = d1;
        part_b[((size_t)tile_id * COUT + c) * 2 + 1] = d2;
      }
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (t == 0) {
        const unsigned int prev = atomicAdd(a.ticket + b * XFB_TICKET_STRIDE, 1u);
        s_last = (prev == (unsigned int)(tiles - 1)) ? 1u : 0u;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (s_last) {
        __threadfence();
        // last CTA of the frame: fixed-order fold of all tile partials (slice-strided, then slice order)
        for (int e = t; e < NSLICE * COUT; e += CTC_PROD) {
          const int c = e % COUT, sl = e / COUT;
          double d1 = 0.0, d2 = 0.0;
          for (int i = sl; i < tiles; i += NSLICE) {
            d1 += __ldcg(part_b + ((size_t)i * COUT + c) * 2);
            d2 += __ldcg(part_b + ((size_t)i * COUT + c) * 2 + 1);
          }
          sStat[(sl * COUT + c) * 2] = d1;
          sStat[(sl * COUT + c) * 2 + 1] = d2;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const double n = (double)a.Hout * (double)a.Wout;
        for (int c = t; c < COUT; c += CTC_PROD) {
          double d1 = 0.0, d2 = 0.0;
          for (int sl = 0; sl < NSLICE; ++sl) { d1 += sStat[(sl * COUT + c) * 2]; d2 += sStat[(sl * COUT + c) * 2 + 1]; }
          const double mean = d1 / n;
          double var = d2 / n - mean * mean;
          if (var < 0.0) var = 0.0;
          a.out_mean[b * COUT + c] = (float)mean;
          a.out_rstd[b * COUT + c] = (float)(1.0 / sqrt(var + 1e-5));
        }
        if (t == 0) a.ticket[b * XFB_TICKET_STRIDE] = 0u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
//                       CIN COUT KS S CSTAGE
using TcB2x = TcCfg<24, 24, 3, 1, 24>;      // block2.0/.1                        120x160
using TcB30 = TcCfg<24, 64, 3, 2, 8>;       // block3.0                           -> 60x80   (3 channel phases: 56 KB smem, 4 CTAs / SM)
using TcC33 = TcCfg<64, 64, 3, 1, 32>;      // block3.1, block4.1/.2, block_fusion.0/.1      (2 channel phases: 98 KB smem, 2 CTAs / SM)
using TcC11 = TcCfg<64, 64, 1, 1, 64>;      // block3.2, block_fusion.2, heatmap_head.0/.1, keypoint_head.0/.1/.2
using TcB40 = TcCfg<64, 64, 3, 2, 16>;      // block4.0                           -> 30x40   (4 channel phases: 107 KB smem, 2 CTAs / SM)
using TcB50 = TcCfg<64, 128, 3, 2, 32>;     // block5.0                           -> 15x20
using TcB5x = TcCfg<128, 128, 3, 1, 64>;    // block5.1/.2
using TcB53 = TcCfg<128, 64, 1, 1, 128>;    // block5.3
using TcKP3 = TcCfg<64, 80, 1, 1, 64>;      // keypoint_head.3 (65 outputs, N padded to 80) + softmax / fold epilogue

struct TcLayerInfo { int cstage, np; };
static TcLayerInfo tc_info(int L) {
  switch (L) {
    case L_B2_0: case L_B2_1: return {TcB2x::CSTAGE, TcB2x::NP};
    case L_B3_0: return {TcB30::CSTAGE, TcB30::NP};
    case L_B4_0: return {TcB40::CSTAGE, TcB40::NP};
    case L_B5_0: return {TcB50::CSTAGE, TcB50::NP};
    case L_B5_1: case L_B5_2: return {TcB5x::CSTAGE, TcB5x::NP};
    case L_B5_3: return {TcB53::CSTAGE, TcB53::NP};
    case L_KP_3: return {TcKP3::CSTAGE, TcKP3::NP};
    case L_B3_1: case L_B4_1: case L_B4_2: case L_F_0: case L_F_1: return {TcC33::CSTAGE, TcC33::NP};
    default: return {TcC11::CSTAGE, TcC11::NP};
  }
}

template <class C, int INMODE, int OUTMODE>
static cudaError_t run_tc(Ctx* c, const ConvTcArgs& a, int tag) {
  auto kern = conv_tc_kernel<C, INMODE, OUTMODE>;
  static unsigned long long attr_mask = 0;
  if (!((attr_mask >> c->device) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << c->device;
  }
  dim3 grid((a.Wout + C::TW - 1) / C::TW, (a.Hout + C::TH - 1) / C::TH, c->B);
  prof_begin(c, tag);
  kern<<<grid, CTC_THREADS, C::SMEM_BYTES, c->stream>>>(a);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

size_t conv_tc_part_elems(int H, int W) {
  size_t m = 0;
  auto upd = [&](int lvl, int cout) {
    const int h = H >> lvl, w = W >> lvl;
    const size_t v = (size_t)((w + 7) / 8) * ((h + 15) / 16) * cout * 2;
    if (v > m) m = v;
  };
  upd(2, 24); upd(3, 64); upd(4, 64); upd(5, 128);
  return m;
}

bool conv_tc_handles(int L) {
  switch (L) {
    case L_B2_0: case L_B2_1: case L_B3_0: case L_B3_1: case L_B3_2: case L_B4_0: case L_B4_1: case L_B4_2: case L_B5_0: case L_B5_1:
    case L_B5_2: case L_B5_3: case L_F_0: case L_F_1: case L_F_2: case L_HM_0: case L_HM_1: case L_KP_0: case L_KP_1: case L_KP_2:
    case L_KP_3:
      return true;
    default:
      return false;
  }
}

// Host side of xfb_create: OIHW fp32 weights -> [phase][tap][hi | lo][cstage/4][np/8][8][4] UMMA operand images
void conv_tc_pack_weights(int L, const float* oihw, int cout, int cin, int ks, std::vector<float>& img) {
  const TcLayerInfo li = tc_info(L);
  const int taps = ks * ks, nphase = cin / li.cstage;
  const size_t unit = (size_t)2 * li.cstage * li.np;
  img.assign((size_t)nphase * taps * unit, 0.f);
  for (int tap = 0; tap < taps; ++tap)
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci) {
        const float w = oihw[((size_t)co * cin + ci) * taps + tap];
        uint32_t u;
        std::memcpy(&u, &w, 4);
        u &= 0xFFFFE000u;
        float hi;
        std::memcpy(&hi, &u, 4);
        const float lo = w - hi;
        const int ph = ci / li.cstage, cil = ci % li.cstage;
        const size_t idx = (size_t)((cil >> 2) * (li.np >> 3) + (co >> 3)) * 32 + (co & 7) * 4 + (cil & 3);
        const size_t base = ((size_t)ph * taps + tap) * unit;
        img[base + idx] = hi;
        img[base + (size_t)li.cstage * li.np + idx] = lo;
      }
}

cudaError_t launch_conv_tc_layer(Ctx* c, int L) {
  const LayerSpec& sp = kLayers[L];
  ConvTcArgs a = {};
  a.Hin = c->H >> sp.lvl_in; a.Win = c->W >> sp.lvl_in;
  a.Hout = c->H >> sp.lvl_out; a.Wout = c->W >> sp.lvl_out;
  a.wimg = c->wimg[L];
  a.bias = c->bias[L];
  a.out = c->act[L];
  a.part = c->part; a.ticket = c->ticket;
  a.full_w = c->W;
  if (L < L_NUM_BN) { a.out_mean = c->bn[L].mean; a.out_rstd = c->bn[L].rstd; }
  auto from = [&](int P) { a.in = c->act[P]; a.in_mean = c->bn[P].mean; a.in_rstd = c->bn[P].rstd; };
  switch (L) {
    case L_B2_0:
      from(L_B1_3);
      a.skip_avg = c->avg4; a.skip_w = c->w[L_SKIP]; a.skip_b = c->bias[L_SKIP];
      return run_tc<TcB2x, TIN_BN_SKIP, TOUT_STATS>(c, a, L);
    case L_B2_1: from(L_B2_0); return run_tc<TcB2x, TIN_BN, TOUT_STATS>(c, a, L);
    case L_B3_0: from(L_B2_1); return run_tc<TcB30, TIN_BN, TOUT_STATS>(c, a, L);
    case L_B3_1: from(L_B3_0); return run_tc<TcC33, TIN_BN, TOUT_STATS>(c, a, L);
    case L_B3_2: from(L_B3_1); return run_tc<TcC11, TIN_BN, TOUT_STATS>(c, a, L);
    case L_B4_0: from(L_B3_2); return run_tc<TcB40, TIN_BN, TOUT_STATS>(c, a, L);
    case L_B4_1: from(L_B4_0); return run_tc<TcC33, TIN_BN, TOUT_STATS>(c, a, L);
    case L_B4_2: from(L_B4_1); return run_tc<TcC33, TIN_BN, TOUT_STATS>(c, a, L);
    case L_B5_0: from(L_B4_2); return run_tc<TcB50, TIN_BN, TOUT_STATS>(c, a, L);
    case L_B5_1: from(L_B5_0); return run_tc<TcB5x, TIN_BN, TOUT_STATS>(c, a, L);
    case L_B5_2: from(L_B5_1); return run_tc<TcB5x, TIN_BN, TOUT_STATS>(c, a, L);
    case L_B5_3: from(L_B5_2); return run_tc<TcB53, TIN_BN, TOUT_STATS>(c, a, L);
    case L_F_0: a.in = c->pyr; return run_tc<TcC33, TIN_PLAIN, TOUT_STATS>(c, a, L);
    case L_F_1: from(L_F_0); return run_tc<TcC33, TIN_BN, TOUT_STATS>(c, a, L);
    case L_F_2: from(L_F_1); return run_tc<TcC11, TIN_BN, TOUT_BIAS>(c, a, L);
    case L_HM_0: a.in = c->act[L_F_2]; return run_tc<TcC11, TIN_PLAIN, TOUT_STATS>(c, a, L);
    case L_HM_1: from(L_HM_0); return run_tc<TcC11, TIN_BN, TOUT_STATS>(c, a, L);
    case L_KP_0: a.in = c->xn; a.Hin = c->H >> 3; a.Win = c->W >> 3; return run_tc<TcC11, TIN_UNFOLD, TOUT_STATS>(c, a, L);
    case L_KP_1: from(L_KP_0); return run_tc<TcC11, TIN_BN, TOUT_STATS>(c, a, L);
    case L_KP_2: from(L_KP_1); return run_tc<TcC11, TIN_BN, TOUT_STATS>(c, a, L);
    case L_KP_3: from(L_KP_2); a.out = c->k1h; return run_tc<TcKP3, TIN_BN, TOUT_KPSOFTMAX>(c, a, L);   // -> K1h [B, H, W]
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace xfb
