// conv_tc2.cu -- the dense convolutions of XFeat (Cin >= 24) as PERSISTENT, PIPELINED implicit GEMMs on tcgen05.
//
// Contract (unchanged from conv_tc.cu; reference: BasicLayerImpl, src/XFeat.cc:7-28 -- conv, TRAIN-mode BatchNorm, ReLU):
// the producer's BN + ReLU (+ skip1 add, + unfold2d) is applied while the input tile is staged, the raw conv output is
// written NHWC fp32, and per-(frame, channel) sum / sum-of-squares go through a fixed-order fold, so a batch of B frames
// equals B batch-1 calls bit for bit.
//
// What is new (round 2):
//   * fp16 two-piece split instead of 3xTF32.  x*16 = hi + lo (hi = fp16 RN, lo = fp16 RN of the exact remainder), weights
//     likewise after an exact power-of-two scale per layer; D += hi*hi + hi*lo + lo*hi with kind::f16 (K = 16 per UTCHMMA).
//     Same ~2^-22 relative operand precision as the TF32 split, at HALF the tensor-pipe time (3 fp16 MMA passes instead
//     of 3 TF32 passes at half rate) and HALF the shared-memory bytes (2 + 2 instead of 4 + 4 bytes per element).
//   * Persistent CTAs (one per SM) looping over 128-pixel output tiles; the layer's weights are loaded into shared memory
//     ONCE per CTA and stay resident (layers whose weight image exceeds shared memory split their output channels over
//     CTAs: COUT_SPLIT); three pipeline stages run concurrently on different tiles:
//        7 producer warps   stage tile i+1 (LDG.256 -> BN/ReLU -> hi/lo split -> STS.128) into a ring of NBUF buffers,
//        1 MMA warp         multiplies tile i   (one elected lane; accumulator = TMEM stage i & 1),
//        4 epilogue warps   drain tile i-1      (tcgen05.ld -> scale -> STG.256 + per-channel statistics).
//   * No im2col, as before: the halo tile is staged as channel-chunk planes [cin/8][pixel][8 halfs], which IS the canonical
//     K-major no-swizzle UMMA layout, so filter tap (ky, kx) is the same buffer with the descriptor start address advanced
//     by (ky * tile_width + kx) * 16 B; stride-2 layers stage four parity planes.  1x1 layers tile the frame linearly.
//   * Statistics without a shared-memory staging tile: each epilogue lane owns one pixel (TMEM lane) and 32 channels; a
//     31-shuffle transpose-reduce leaves the sum of channel j in lane j.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "tc_ptx.cuh"
#include "xfb_internal.h"

namespace xfb {

enum Tc2InMode { T2IN_PLAIN = 0, T2IN_BN = 1, T2IN_BN_SKIP = 2, T2IN_UNFOLD = 3 };
enum Tc2OutMode { T2OUT_STATS = 0, T2OUT_BIAS = 1, T2OUT_KPSOFTMAX = 2 };

constexpr float T2_ACT_SCALE = 16.0f;        // activations are staged as x * 16 (exact), |x| < 4094 fits fp16
constexpr int T2_PROD_WARPS = 7;           // 12 warps in all: the register file is allocated per 4 warps, so 13 warps would cap a thread at 128 registers
constexpr int T2_PROD = 32 * T2_PROD_WARPS;  // producer threads
constexpr int T2_EPI = 128;                  // epilogue threads (warps 0-3: TMEM lane quadrant = warp id)
constexpr int T2_THREADS = T2_EPI + 32 + T2_PROD;   // warps 0-3 epilogue, warp 4 MMA issuer, warps 5-11 producers
constexpr int T2_UNR = 4;                    // independent 32-byte global loads in flight per producer thread

// CIN: real input channels; CINP: padded to a multiple of 16 (zero chunks); COUT: real outputs; NSPLIT: output-channel
// groups processed by different CTAs; CSTAGE: input channels per staged buffer (channel phase); NBUF: ring depth.
template <int CIN_, int COUT_, int KS_, int S_, int CSTAGE_, int NBUF_, int NSPLIT_>
struct T2Cfg {
  static constexpr int CIN = CIN_, COUT = COUT_, KS = KS_, S = S_, CSTAGE = CSTAGE_, NBUF = NBUF_, NSPLIT = NSPLIT_;
  static constexpr int CINP = (CIN + 15) / 16 * 16;
  static constexpr int NOUT = COUT / NSPLIT;                                      // real output channels per CTA
  static constexpr int NP = NOUT <= 32 ? 32 : (NOUT + 15) / 16 * 16;              // UMMA N (M = 128 needs N % 16 == 0)
  static constexpr int ACC_STRIDE = NP <= 32 ? 32 : (NP <= 64 ? 64 : 128);        // TMEM columns per accumulator stage
  static constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int TW = 8, TH = 16, PAD = KS / 2;
  static constexpr int NSUB = S == 2 ? 4 : 1;                                     // parity planes
  static constexpr int WT = S == 2 ? TW + 1 : TW + 2 * PAD;
  static constexpr int HT = S == 2 ? TH + 1 : TH + 2 * PAD;
  static constexpr int NPIX = WT * HT;
  static constexpr int KCS = CSTAGE / 8;                                          // 16-byte channel chunks (8 halfs) per phase
  // plane stride (bytes): pixels * 16, padded so that the 8 lanes of a quarter warp (kc fastest) hit distinct 16-byte bank groups
  static constexpr int PS_RAW = NPIX * 16;
  static constexpr int PS_MOD = KCS >= 8 ? 16 : 128 / KCS;
  static constexpr int PS = PS_RAW + ((PS_MOD - PS_RAW % 128) % 128 + 128) % 128;
  static constexpr int SUB_BYTES = KCS * PS;
  static constexpr int IN_BYTES = NSUB * SUB_BYTES;                               // one of hi / lo
  static constexpr int BUF_BYTES = 2 * IN_BYTES;
  static constexpr int TAPS = KS * KS;
  static constexpr int NPHASE = CINP / CSTAGE;
  static constexpr int W_UNIT_BYTES = CSTAGE * NP * 2;                            // one of hi / lo of one (phase, tap)
  static constexpr int W_BYTES = NPHASE * TAPS * 2 * W_UNIT_BYTES;                // resident weight image of one output group
  static constexpr uint32_t LBO_A = PS, SBO_A = WT * 16;
  static constexpr uint32_t LBO_B = (NP / 8) * 128, SBO_B = 128;
  static constexpr int STAT_BYTES = 4 * 64 * 2 * 4 + 128 * 2 * 8;                 // quadrant sums [4][64][2] (float) + fold scratch [128][2] (double)
  static constexpr size_t SMEM_BYTES = (size_t)W_BYTES + (size_t)NBUF * BUF_BYTES + STAT_BYTES + 256;
  static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((128u >> 4) << 24);   // kind::f16, F16 x F16 -> F32, K-major
  static_assert(CSTAGE % 16 == 0 && CINP % CSTAGE == 0 && COUT % NSPLIT == 0 && (NOUT % 8 == 0 || NOUT == 65) && NOUT <= 80, "shape");
  static_assert(S == 1 || KS == 3, "stride 2 is implemented for 3x3 only");
  static_assert(T2_PROD % KCS == 0, "a producer thread keeps one channel chunk");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(PS % 16 == 0 && W_UNIT_BYTES % 16 == 0, "descriptor alignment");
};

struct ConvTc2Args {
  const float* in;          // NHWC raw producer output (or plain values), or xn for T2IN_UNFOLD
  const unsigned char* wimg;   // [split][phase][tap][hi | lo][cstage/8][np/8][8][8 halfs]
  const float* bias;
  float* out;
  const float* in_mean; const float* in_rstd;
  const float* skip_avg; const float* skip_w; const float* skip_b;
  float* part;              // per-tile channel sums [B][tiles][COUT][2] (float)
  unsigned int* ticket; float* out_mean; float* out_rstd;
  int B, Hin, Win, Hout, Wout;
  int full_w;               // T2IN_UNFOLD / KPSOFTMAX: width of xn / K1h
  int tiles_x, tiles;       // tiles per frame
  float out_scale;          // 1 / (activation scale * weight scale), an exact power of two
  unsigned long long* dbg;  // XFB_T2_DEBUG: cycle counters of CTA 0 (32 per layer), else nullptr
};

// debug: add the cycles since `t0` to counter `slot` (one designated thread per role) and restart the clock
#define T2_TICK(slot)                                                                  \
  do {                                                                                 \
    if (dbg_me) { const long long _n = clock64(); atomicAdd(a.dbg + (slot), (unsigned long long)(_n - t0)); t0 = _n; } \
  } while (0)

struct f8 { float4 a, b; };
__device__ __forceinline__ f8 ldg256(const float* p) {
  f8 r;
  asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f), "f"(g), "f"(h) : "memory");
}
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// x (already scaled) -> fp16 hi (round to nearest) and fp16 lo = RN(x - hi); two values per call
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
                 "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_nw(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
      "%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
        "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]),
        "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]),
        "=r"(u[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_fence() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Sum over the 32 lanes of v[j] for every j: afterwards lane j holds the total of column j.  31 shuffles.
__device__ __forceinline__ float transpose_reduce32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float keep = up ? v[i + off] : v[i];
      const float send = up ? v[i] : v[i + off];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <class C, int INMODE, int OUTMODE>
__global__ void __launch_bounds__(T2_THREADS, 1) conv_tc2_kernel(const ConvTc2Args a) {
  constexpr int KCS = C::KCS, PS = C::PS, WT = C::WT, NPIX = C::NPIX, TAPS = C::TAPS, NBUF = C::NBUF, NPHASE = C::NPHASE, NSUB = C::NSUB;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sW = smem_raw;                                   // resident weights of this CTA's output group
  unsigned char* sA = smem_raw + C::W_BYTES;                      // NBUF x [hi | lo] staged tiles
  float* sQ = reinterpret_cast<float*>(sA + (size_t)NBUF * C::BUF_BYTES);     // [4][64][2] quadrant sums
  double* sFold = reinterpret_cast<double*>(sQ + 4 * 64 * 2);                  // [128][2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sFold + 128 * 2);
  uint64_t* bar_in = bars + 0;          // [NBUF] tile staged (T2_PROD arrivals)
  uint64_t* bar_free = bars + 4;        // [NBUF] tensor core finished reading the staged buffer
  uint64_t* bar_accf = bars + 8;        // [2] accumulator complete
  uint64_t* bar_acce = bars + 10;       // [2] accumulator drained (4 arrivals)
  uint64_t* bar_w = bars + 12;          // [NPHASE <= 8] weights of a channel phase landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 20);
  uint32_t* s_flag = s_tmem + 1;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long long t_kernel0 = clock64();
  // work items: (tile of a frame, output group).  A CTA keeps ONE output group (its weights stay resident) and walks the
  // tiles of all frames with stride gridDim.x / NSPLIT.
  const int split = blockIdx.x % C::NSPLIT;
  const int first = blockIdx.x / C::NSPLIT, stride = gridDim.x / C::NSPLIT;
  const int n_tiles = a.B * a.tiles;

  if (t == 0) {
    for (int s = 0; s < NBUF; ++s) { mbar_init(bar_in + s, T2_PROD); mbar_init(bar_free + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_accf + s, 1); mbar_init(bar_acce + s, 4); }
    for (int s = 0; s < NPHASE; ++s) mbar_init(bar_w + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(s_tmem, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp >= 5) {
    // ===================== producers: stage (tile, channel phase) buffers =====================
    const int pt = t - (T2_EPI + 32);
    const int kc = pt % KCS;                         // this thread's 8-channel chunk inside a phase (constant)
    constexpr int ITEMS = NSUB * NPIX * KCS;
    int it = 0;
    const bool dbg_me = a.dbg != nullptr && blockIdx.x == 0 && pt == 0;
    long long t0 = dbg_me ? clock64() : 0;
    for (int tile = first; tile < n_tiles; tile += stride) {
      const int b = tile / a.tiles, tf = tile - b * a.tiles;
      const int oy0 = (C::KS == 1) ? 0 : (tf / a.tiles_x) * C::TH, ox0 = (C::KS == 1) ? 0 : (tf % a.tiles_x) * C::TW;
      const float* in_b = (INMODE == T2IN_UNFOLD) ? a.in + (size_t)b * (a.Hin * 8) * a.full_w : a.in + (size_t)b * a.Hin * a.Win * C::CIN;
      for (int ph = 0; ph < NPHASE; ++ph, ++it) {
        const int buf = it % NBUF;
        unsigned char* dst_hi = sA + (size_t)buf * C::BUF_BYTES;
        const int ch = ph * C::CSTAGE + kc * 8;      // first of this thread's 8 channels
        const bool ch_ok = ch < C::CIN;              // padded chunks (Cin 24 -> 32) are zero
        float m[8], r[8], sw[8], sb[8];
        if ((INMODE == T2IN_BN || INMODE == T2IN_BN_SKIP) && ch_ok) {
          const f8 mm = ldg256(a.in_mean + b * C::CIN + ch), rr = ldg256(a.in_rstd + b * C::CIN + ch);
          m[0] = mm.a.x; m[1] = mm.a.y; m[2] = mm.a.z; m[3] = mm.a.w; m[4] = mm.b.x; m[5] = mm.b.y; m[6] = mm.b.z; m[7] = mm.b.w;
          r[0] = rr.a.x; r[1] = rr.a.y; r[2] = rr.a.z; r[3] = rr.a.w; r[4] = rr.b.x; r[5] = rr.b.y; r[6] = rr.b.z; r[7] = rr.b.w;
          if (INMODE == T2IN_BN_SKIP) {
            const f8 ww = ldg256(a.skip_w + ch), bb = ldg256(a.skip_b + ch);
            sw[0] = ww.a.x; sw[1] = ww.a.y; sw[2] = ww.a.z; sw[3] = ww.a.w; sw[4] = ww.b.x; sw[5] = ww.b.y; sw[6] = ww.b.z; sw[7] = ww.b.w;
            sb[0] = bb.a.x; sb[1] = bb.a.y; sb[2] = bb.a.z; sb[3] = bb.a.w; sb[4] = bb.b.x; sb[5] = bb.b.y; sb[6] = bb.b.z; sb[7] = bb.b.w;
          }
        }
        T2_TICK(3);
        if (it >= NBUF) mbar_wait(bar_free + buf, ((it / NBUF) - 1) & 1);      // the MMAs that read this buffer are done
        T2_TICK(2);
        for (int base = 0; base < ITEMS; base += T2_PROD * T2_UNR) {
          f8 v[T2_UNR];
          float av[T2_UNR];
          int off[T2_UNR];
          bool inside[T2_UNR];
          // all loads of the batch first (memory-level parallelism), then transform and store
#pragma unroll
          for (int u = 0; u < T2_UNR; ++u) {
            const int idx = base + u * T2_PROD + pt;
            v[u].a = make_float4(0.f, 0.f, 0.f, 0.f); v[u].b = v[u].a;
            av[u] = 0.f; off[u] = -1; inside[u] = false;
            if (idx < ITEMS) {
              const int rest = idx / KCS;            // idx % KCS == kc
              const int pix = rest % NPIX, sub = rest / NPIX;
              int iy, ix;
              if (C::KS == 1) { const int lin = tf * 128 + pix; iy = lin / a.Win; ix = lin - iy * a.Win; }
              else if (C::S == 1) { iy = oy0 - C::PAD + pix / WT; ix = ox0 - C::PAD + pix % WT; }
              else { iy = 2 * (oy0 - 1 + pix / WT) + (sub >> 1); ix = 2 * (ox0 - 1 + pix % WT) + (sub & 1); }
              off[u] = sub * C::SUB_BYTES + kc * PS + pix * 16;
              if (ch_ok && iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) {
                inside[u] = true;
                if (INMODE == T2IN_UNFOLD) {
                  // XFeatModel::unfold2d(x, 8), src/XFeat.cc:124-133: channel c = (y % 8) * 8 + x % 8 -> chunk kc = row iy * 8 + kc of xn
                  v[u] = ldg256(in_b + (size_t)(iy * 8 + (ch >> 3)) * a.full_w + ix * 8);
                } else {
                  v[u] = ldg256(in_b + ((size_t)iy * a.Win + ix) * C::CIN + ch);
                  if (INMODE == T2IN_BN_SKIP) av[u] = a.skip_avg[((size_t)b * a.Hin + iy) * a.Win + ix];
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < T2_UNR; ++u) {
            if (off[u] < 0) continue;
            float x[8] = {v[u].a.x, v[u].a.y, v[u].a.z, v[u].a.w, v[u].b.x, v[u].b.y, v[u].b.z, v[u].b.w};
            if (inside[u]) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                if (INMODE == T2IN_BN || INMODE == T2IN_BN_SKIP) {
                  x[e] = fmaxf((x[e] - m[e]) * r[e], 0.f);
                  // x1 + skip1(x), src/XFeat.cc:153; skip1 = AvgPool2d(4,4) + Conv2d(1,24,1) (:36-39)
                  if (INMODE == T2IN_BN_SKIP) x[e] += av[u] * sw[e] + sb[e];
                }
                x[e] *= T2_ACT_SCALE;
              }
            }
            uint4 hi, lo;
            split2(x[0], x[1], hi.x, lo.x); split2(x[2], x[3], hi.y, lo.y); split2(x[4], x[5], hi.z, lo.z); split2(x[6], x[7], hi.w, lo.w);
            *reinterpret_cast<uint4*>(dst_hi + off[u]) = hi;
            *reinterpret_cast<uint4*>(dst_hi + C::IN_BYTES + off[u]) = lo;
          }
        }
        T2_TICK(3);
        fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
        mbar_arrive(bar_in + buf);
        T2_TICK(4);
        if (dbg_me) atomicAdd(a.dbg + 5, 1ull);
      }
    }
  } else if (warp == 4) {
    // ===================== weights (once) + MMA issue: the whole warp runs the uniform loop, ONE elected lane issues ===========
    const bool leader = elect_one_sync();
    if (leader && first < n_tiles) {
      const unsigned char* wsrc = a.wimg + (size_t)split * C::W_BYTES;
      constexpr uint32_t PH_BYTES = (uint32_t)TAPS * 2 * C::W_UNIT_BYTES;
      for (int ph = 0; ph < NPHASE; ++ph) {
        mbar_expect_tx(bar_w + ph, PH_BYTES);
        for (int u = 0; u < TAPS * 2; ++u)
          bulk_g2s(sW + (size_t)ph * PH_BYTES + (size_t)u * C::W_UNIT_BYTES, wsrc + (size_t)ph * PH_BYTES + (size_t)u * C::W_UNIT_BYTES, C::W_UNIT_BYTES,
                   bar_w + ph);
      }
    }
    __syncwarp();
    const uint64_t da0 = umma_desc_kmajor(smem_u32(sA), C::LBO_A, C::SBO_A);
    const uint64_t db0 = umma_desc_kmajor(smem_u32(sW), C::LBO_B, C::SBO_B);
    constexpr uint64_t KA = (2u * C::LBO_A) >> 4, KB = (2u * C::LBO_B) >> 4;      // one K = 16 step (two 16-byte chunks)
    constexpr uint64_t A_LO = (uint64_t)C::IN_BYTES >> 4, A_BUF = (uint64_t)C::BUF_BYTES >> 4;
    constexpr uint64_t W_UNIT = (uint64_t)C::W_UNIT_BYTES >> 4;
    int it = 0, n = 0;
    const bool dbg_me = a.dbg != nullptr && blockIdx.x == 0 && leader;
    long long t0 = dbg_me ? clock64() : 0;
    for (int tile = first; tile < n_tiles; tile += stride, ++n) {
      const int acc = n & 1;
      T2_TICK(8);
      if (n >= 2) mbar_wait(bar_acce + acc, ((n >> 1) - 1) & 1);                 // the epilogue has drained this accumulator
      T2_TICK(7);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)acc * C::ACC_STRIDE;
      for (int ph = 0; ph < NPHASE; ++ph, ++it) {
        const int buf = it % NBUF;
        T2_TICK(8);
        if (n == 0) mbar_wait(bar_w + ph, 0);
        T2_TICK(9);
        mbar_wait(bar_in + buf, (it / NBUF) & 1);
        T2_TICK(6);
        tc_fence_after();
        if (leader) {
          const uint64_t dah = da0 + (uint64_t)buf * A_BUF, dal = dah + A_LO;
#pragma unroll
          for (int tap = 0; tap < TAPS; ++tap) {
            // shifted window of the staged tile (stride 2: the parity plane of this tap)
            const int ky = tap / C::KS, kx = tap % C::KS;
            uint32_t tap_off;
            if (C::S == 1) tap_off = (uint32_t)(ky * WT + kx) * 16u;
            else {
              const int py = (ky == 1) ? 0 : 1, dy = (ky == 0) ? 0 : 1, px = (kx == 1) ? 0 : 1, dx = (kx == 0) ? 0 : 1;
              tap_off = (uint32_t)((py * 2 + px) * C::SUB_BYTES) + (uint32_t)(dy * WT + dx) * 16u;
            }
            const uint64_t ah = dah + (tap_off >> 4), al = dal + (tap_off >> 4);
            const uint64_t bh = db0 + (uint64_t)((ph * TAPS + tap) * 2) * W_UNIT, bl = bh + W_UNIT;
#pragma unroll
            for (int k16 = 0; k16 < C::CSTAGE / 16; ++k16) {
              umma_f16_ss(d, ah + k16 * KA, bh + k16 * KB, C::IDESC, (ph > 0 || tap > 0 || k16 > 0) ? 1u : 0u);
              umma_f16_ss(d, ah + k16 * KA, bl + k16 * KB, C::IDESC, 1u);
              umma_f16_ss(d, al + k16 * KA, bh + k16 * KB, C::IDESC, 1u);
            }
          }
          umma_commit(bar_free + buf);                // the staged buffer may be refilled once these MMAs have read it
          if (ph == NPHASE - 1) umma_commit(bar_accf + acc);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global (+ statistics) =====================
    const int quad = warp;                              // TMEM lane quadrant (hardware: warp id % 4)
    const int p = quad * 32 + lane;                     // pixel of the tile = TMEM lane
    const int cbase = split * C::NOUT;                  // first output channel of this CTA's group
    int n = 0;
    const bool dbg_me = a.dbg != nullptr && blockIdx.x == 0 && t == 0;
    long long t0 = dbg_me ? clock64() : 0;
    for (int tile = first; tile < n_tiles; tile += stride, ++n) {
      const int acc = n & 1;
      const int b = tile / a.tiles, tf = tile - b * a.tiles;
      if (dbg_me) atomicAdd(a.dbg + 16, 1ull);
      int oy, ox;
      if (C::KS == 1) { const int lin = tf * 128 + p; oy = lin / a.Wout; ox = lin - oy * a.Wout; }
      else { oy = (tf / a.tiles_x) * C::TH + (p >> 3); ox = (tf % a.tiles_x) * C::TW + (p & 7); }
      const bool valid = oy < a.Hout && ox < a.Wout;
      const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * C::ACC_STRIDE;
      T2_TICK(15);
      mbar_wait(bar_accf + acc, (n >> 1) & 1);
      tc_fence_after();
      T2_TICK(10);
      if constexpr (OUTMODE == T2OUT_KPSOFTMAX) {
        // keypoint_head.3 epilogue: + bias, softmax over the 65 logits, drop the dustbin, 8x8 fold
        // (src/XFeat.cc:85-90, XFextractor::getKptsHeatmap src/XFextractor.cc:204-217); one thread per cell
        float v[80];
        tmem_ld32_nw(tq, v); tmem_ld32_nw(tq + 32u, v + 32); tmem_ld16(tq + 64u, v + 64);
        tmem_ld_fence();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acce + acc);
        float mx = -3.4e38f;
#pragma unroll
        for (int c = 0; c < 65; ++c) { v[c] = fmaf(v[c], a.out_scale, a.bias[c]); mx = fmaxf(mx, v[c]); }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 65; ++c) { v[c] = expf(v[c] - mx); sum += v[c]; }
        if (valid) {
          float* dst = a.out + (size_t)b * (a.Hout * 8) * a.full_w + (size_t)(oy * 8) * a.full_w + ox * 8;
#pragma unroll
          for (int ry = 0; ry < 8; ++ry)
            stg256(dst + (size_t)ry * a.full_w, v[ry * 8 + 0] / sum, v[ry * 8 + 1] / sum, v[ry * 8 + 2] / sum, v[ry * 8 + 3] / sum, v[ry * 8 + 4] / sum,
                   v[ry * 8 + 5] / sum, v[ry * 8 + 6] / sum, v[ry * 8 + 7] / sum);
        }
      } else {
        constexpr int NCH = C::NP / 32;                 // 32-column chunks (NP is 32 or 64 here)
        static_assert(C::NP % 32 == 0, "epilogue chunking");
        float v[NCH][32];
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) tmem_ld32_nw(tq + (uint32_t)ci * 32u, v[ci]);
        tmem_ld_fence();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acce + acc);     // values are in registers: the tensor core may reuse the accumulator
        float* orow = a.out + (((size_t)b * a.Hout + oy) * a.Wout + ox) * C::COUT + cbase;
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            float x = v[ci][q] * a.out_scale;
            if (OUTMODE == T2OUT_BIAS && ci * 32 + q < C::NOUT) x += a.bias[cbase + ci * 32 + q];
            v[ci][q] = x;
          }
          if (valid) {
#pragma unroll
            for (int q = 0; q < 32; q += 8)
              if (ci * 32 + q < C::NOUT)
                stg256(orow + ci * 32 + q, v[ci][q], v[ci][q + 1], v[ci][q + 2], v[ci][q + 3], v[ci][q + 4], v[ci][q + 5], v[ci][q + 6], v[ci][q + 7]);
          }
        }
        T2_TICK(11);
        if constexpr (OUTMODE == T2OUT_STATS) {
          // per-channel sum / sum of squares over the valid pixels of this warp's quadrant: lane j <- channel ci * 32 + j
#pragma unroll
          for (int ci = 0; ci < NCH; ++ci) {
            float sq[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) { const float x = valid ? v[ci][q] : 0.f; v[ci][q] = x; sq[q] = x * x; }
            const float s1 = transpose_reduce32(v[ci], lane);
            const float s2 = transpose_reduce32(sq, lane);
            *reinterpret_cast<float2*>(sQ + ((quad * 64) + ci * 32 + lane) * 2) = make_float2(s1, s2);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          T2_TICK(12);
          // tile partial = the four quadrants in fixed order -> part[b][tile][channel]
          float* part_b = a.part + (size_t)b * a.tiles * C::COUT * 2;
          if (t < C::NOUT) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) { const float2 x = *reinterpret_cast<const float2*>(sQ + ((qd * 64) + t) * 2); s1 += x.x; s2 += x.y; }
            *reinterpret_cast<float2*>(part_b + ((size_t)tf * C::COUT + cbase + t) * 2) = make_float2(s1, s2);
          }
          __threadfence();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          T2_TICK(13);
          if (t == 0) {
            const unsigned int prev = atomicAdd(a.ticket + b * XFB_TICKET_STRIDE, 1u);
            *s_flag = (prev == (unsigned int)(a.tiles * C::NSPLIT - 1)) ? 1u : 0u;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          T2_TICK(14);
          if (*s_flag) {
            __threadfence();
            // last work item of the frame: fixed-order fold of all tile partials (slice-strided, then slice order), in double
            constexpr int NSL = C::COUT >= 128 ? 1 : (C::COUT >= 64 ? 2 : 4);
            if (t < NSL * C::COUT) {
              const int c = t % C::COUT, sl = t / C::COUT;
              double d1 = 0.0, d2 = 0.0;
              int i = sl;
              for (; i + 7 * NSL < a.tiles; i += 8 * NSL) {        // 8 independent loads in flight, summed in index order
                float2 x[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) x[q] = __ldcg(reinterpret_cast<const float2*>(part_b + ((size_t)(i + q * NSL) * C::COUT + c) * 2));
#pragma unroll
                for (int q = 0; q < 8; ++q) { d1 += (double)x[q].x; d2 += (double)x[q].y; }
              }
              for (; i < a.tiles; i += NSL) {
                const float2 x = __ldcg(reinterpret_cast<const float2*>(part_b + ((size_t)i * C::COUT + c) * 2));
                d1 += (double)x.x; d2 += (double)x.y;
              }
              sFold[t * 2] = d1; sFold[t * 2 + 1] = d2;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (t < C::COUT) {
              double d1 = 0.0, d2 = 0.0;
              for (int sl = 0; sl < NSL; ++sl) { d1 += sFold[(sl * C::COUT + t) * 2]; d2 += sFold[(sl * C::COUT + t) * 2 + 1]; }
              const double cnt = (double)a.Hout * (double)a.Wout;
              const double mean = d1 / cnt;
              double var = d2 / cnt - mean * mean;
              if (var < 0.0) var = 0.0;
              a.out_mean[b * C::COUT + t] = (float)mean;
              a.out_rstd[b * C::COUT + t] = (float)(1.0 / sqrt(var + 1e-5));
            }
            if (t == 0) a.ticket[b * XFB_TICKET_STRIDE] = 0u;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");   // sQ / sFold / s_flag are reused by the next tile
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (a.dbg != nullptr && blockIdx.x == 0 && t == 0) { atomicAdd(a.dbg + 0, 1ull); atomicAdd(a.dbg + 1, (unsigned long long)(clock64() - t_kernel0)); }
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------
//                        CIN COUT KS S CSTAGE NBUF NSPLIT
using T2B2x = T2Cfg<24, 24, 3, 1, 32, 3, 1>;      // block2.0/.1                   120x160         (Cin padded 24 -> 32)
using T2B30 = T2Cfg<24, 64, 3, 2, 16, 3, 1>;      // block3.0                      -> 60x80
using T2C33 = T2Cfg<64, 64, 3, 1, 32, 3, 1>;      // block3.1, block4.1/.2, block_fusion.0/.1   (147 KB of weights resident)
using T2C11 = T2Cfg<64, 64, 1, 1, 64, 3, 1>;      // block3.2, block_fusion.2, heatmap_head.0/.1, keypoint_head.0/.1/.2
using T2B40 = T2Cfg<64, 64, 3, 2, 16, 2, 1>;      // block4.0                      -> 30x40
using T2B50 = T2Cfg<64, 128, 3, 2, 16, 2, 2>;     // block5.0                      -> 15x20        (two output groups of 64)
using T2B5x = T2Cfg<128, 128, 3, 1, 32, 3, 4>;    // block5.1/.2                                   (four output groups of 32)
using T2B53 = T2Cfg<128, 64, 1, 1, 64, 3, 1>;     // block5.3
using T2KP3 = T2Cfg<64, 65, 1, 1, 64, 3, 1>;      // keypoint_head.3 (65 outputs, N padded to 80) + softmax / fold epilogue

struct T2LayerInfo { int cstage, np, nsplit, cinp; };
template <class C> static T2LayerInfo t2_info_of() { return {C::CSTAGE, C::NP, C::NSPLIT, C::CINP}; }
static T2LayerInfo t2_info(int L) {
  switch (L) {
    case L_B2_0: case L_B2_1: return t2_info_of<T2B2x>();
    case L_B3_0: return t2_info_of<T2B30>();
    case L_B4_0: return t2_info_of<T2B40>();
    case L_B5_0: return t2_info_of<T2B50>();
    case L_B5_1: case L_B5_2: return t2_info_of<T2B5x>();
    case L_B5_3: return t2_info_of<T2B53>();
    case L_KP_3: return t2_info_of<T2KP3>();
    case L_B3_1: case L_B4_1: case L_B4_2: case L_F_0: case L_F_1: return t2_info_of<T2C33>();
    default: return t2_info_of<T2C11>();
  }
}

template <class C, int INMODE, int OUTMODE>
static cudaError_t run_tc2(Ctx* c, ConvTc2Args& a, int tag) {
  auto kern = conv_tc2_kernel<C, INMODE, OUTMODE>;
  static unsigned long long attr_mask = 0;
  if (!((attr_mask >> c->device) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << c->device;
  }
  a.B = c->B;
  if (C::KS == 1) { a.tiles_x = 1; a.tiles = (a.Hout * a.Wout + 127) / 128; }
  else { a.tiles_x = (a.Wout + C::TW - 1) / C::TW; a.tiles = a.tiles_x * ((a.Hout + C::TH - 1) / C::TH); }
  const int items = a.B * a.tiles * C::NSPLIT;
  int grid = c->num_sms - c->num_sms % C::NSPLIT;            // persistent: one CTA per SM, a multiple of the output groups
  if (grid > items) grid = items;                            // (items is a multiple of NSPLIT)
  prof_begin(c, tag);
  kern<<<grid, T2_THREADS, C::SMEM_BYTES, c->stream>>>(a);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

// floats of partial-sum scratch per frame: [tiles][COUT][2]
size_t conv_tc2_part_floats(int H, int W) {
  size_t m = 0;
  auto upd = [&](int lvl, int cout, bool linear) {
    const int h = H >> lvl, w = W >> lvl;
    const size_t tiles = linear ? (size_t)(h * w + 127) / 128 : (size_t)((w + 7) / 8) * ((h + 15) / 16);
    const size_t v = tiles * cout * 2;
    if (v > m) m = v;
  };
  upd(2, 24, false); upd(3, 64, false); upd(3, 64, true); upd(4, 64, false); upd(5, 128, false); upd(5, 64, true);
  return m;
}

// Host side of xfb_create: OIHW fp32 weights -> [split][phase][tap][hi | lo][cstage/8][np/8][8][8 halfs] UMMA operand images of
// w * 2^k (k chosen so that max |w| * 2^k is in [1024, 2048)); returns the epilogue factor 1 / (16 * 2^k).
float conv_tc2_pack_weights(int L, const float* oihw, int cout, int cin, int ks, std::vector<unsigned char>& img) {
  const T2LayerInfo li = t2_info(L);
  const int taps = ks * ks, nphase = li.cinp / li.cstage, nout = cout / li.nsplit;
  float wmax = 0.f;
  for (size_t i = 0; i < (size_t)cout * cin * taps; ++i) wmax = std::fmax(wmax, std::fabs(oihw[i]));
  int k = 0;
  if (wmax > 0.f) { int e; std::frexp(wmax, &e); k = 11 - e; }           // wmax = f * 2^e, f in [0.5, 1)  ->  wmax * 2^k in [1024, 2048)
  const float wscale = std::ldexp(1.0f, k);
  const size_t unit = (size_t)li.cstage * li.np * 2;                     // bytes of one hi (or lo) image of a (phase, tap)
  const size_t group = (size_t)nphase * taps * 2 * unit;
  img.assign((size_t)li.nsplit * group, 0);
  for (int tap = 0; tap < taps; ++tap)
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci) {
        const float w = oihw[((size_t)co * cin + ci) * taps + tap] * wscale;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const int sp = co / nout, col = co % nout;
        const int ph = ci / li.cstage, cil = ci % li.cstage;
        const size_t off = ((size_t)(cil >> 3) * (li.np >> 3) + (col >> 3)) * 128 + (size_t)(col & 7) * 16 + (size_t)(cil & 7) * 2;
        const size_t base = (size_t)sp * group + ((size_t)ph * taps + tap) * 2 * unit;
        std::memcpy(&img[base + off], &hi, 2);
        std::memcpy(&img[base + unit + off], &lo, 2);
      }
  return 1.0f / (T2_ACT_SCALE * wscale);
}

cudaError_t launch_conv_tc2_layer(Ctx* c, int L) {
  const LayerSpec& sp = kLayers[L];
  ConvTc2Args a = {};
  a.Hin = c->H >> sp.lvl_in; a.Win = c->W >> sp.lvl_in;
  a.Hout = c->H >> sp.lvl_out; a.Wout = c->W >> sp.lvl_out;
  a.wimg = c->wimg2[L];
  a.out_scale = c->wscale2[L];
  a.bias = c->bias[L];
  a.out = c->act[L];
  a.part = reinterpret_cast<float*>(c->part); a.ticket = c->ticket;
  a.full_w = c->W;
  a.dbg = c->t2_counters ? c->t2_counters + (size_t)L * 32 : nullptr;
  if (L < L_NUM_BN) { a.out_mean = c->bn[L].mean; a.out_rstd = c->bn[L].rstd; }
  auto from = [&](int P) { a.in = c->act[P]; a.in_mean = c->bn[P].mean; a.in_rstd = c->bn[P].rstd; };
  switch (L) {
    case L_B2_0:
      from(L_B1_3);
      a.skip_avg = c->avg4; a.skip_w = c->w[L_SKIP]; a.skip_b = c->bias[L_SKIP];
      return run_tc2<T2B2x, T2IN_BN_SKIP, T2OUT_STATS>(c, a, L);
    case L_B2_1: from(L_B2_0); return run_tc2<T2B2x, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_B3_0: from(L_B2_1); return run_tc2<T2B30, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_B3_1: from(L_B3_0); return run_tc2<T2C33, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_B3_2: from(L_B3_1); return run_tc2<T2C11, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_B4_0: from(L_B3_2); return run_tc2<T2B40, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_B4_1: from(L_B4_0); return run_tc2<T2C33, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_B4_2: from(L_B4_1); return run_tc2<T2C33, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_B5_0: from(L_B4_2); return run_tc2<T2B50, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_B5_1: from(L_B5_0); return run_tc2<T2B5x, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_B5_2: from(L_B5_1); return run_tc2<T2B5x, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_B5_3: from(L_B5_2); return run_tc2<T2B53, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_F_0: a.in = c->pyr; return run_tc2<T2C33, T2IN_PLAIN, T2OUT_STATS>(c, a, L);
    case L_F_1: from(L_F_0); return run_tc2<T2C33, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_F_2: from(L_F_1); return run_tc2<T2C11, T2IN_BN, T2OUT_BIAS>(c, a, L);
    case L_HM_0: a.in = c->act[L_F_2]; return run_tc2<T2C11, T2IN_PLAIN, T2OUT_STATS>(c, a, L);
    case L_HM_1: from(L_HM_0); return run_tc2<T2C11, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_KP_0: a.in = c->xn; a.Hin = c->H >> 3; a.Win = c->W >> 3; return run_tc2<T2C11, T2IN_UNFOLD, T2OUT_STATS>(c, a, L);
    case L_KP_1: from(L_KP_0); return run_tc2<T2C11, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_KP_2: from(L_KP_1); return run_tc2<T2C11, T2IN_BN, T2OUT_STATS>(c, a, L);
    case L_KP_3: from(L_KP_2); a.out = c->k1h; return run_tc2<T2KP3, T2IN_BN, T2OUT_KPSOFTMAX>(c, a, L);   // -> K1h [B, H, W]
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace xfb
