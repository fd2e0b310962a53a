// conv_tc2.cu -- the dense convolutions of XFeat (Cin >= 24) as PERSISTENT, PIPELINED implicit GEMMs on tcgen05.
//
// Contract (unchanged from conv_tc.cu; reference: BasicLayerImpl, src/XFeat.cc:7-28 -- conv, TRAIN-mode BatchNorm, ReLU):
// the producer's BN + ReLU (+ skip1 add, + unfold2d) is applied while the input tile is staged, the raw conv output is
// written NHWC fp32, and per-(frame, channel) sum / sum-of-squares go through a fixed-order fold, so a batch of B frames
// equals B batch-1 calls bit for bit.
//
// Design (round 2):
//   * fp16 two-piece split instead of 3xTF32.  x*16 = hi + lo (hi = fp16 RN, lo = fp16 RN of the exact remainder), weights
//     likewise after an exact power-of-two scale per layer; D = hi*hi + hi*lo + lo*hi with kind::f16 (K = 16 per UTCHMMA).
//     Same ~2^-22 relative operand precision as the TF32 split at half the tensor-pipe time and half the shared-memory
//     bytes.  The hi and lo weight rows are CONCATENATED along N: one MMA computes A_hi x [W_hi ; W_lo] (N = 2*NP, the
//     A operand is fetched from shared memory once for two products), a second one adds A_lo x W_hi into the first NP
//     accumulator columns; the epilogue sums the two column groups.
//   * Persistent CTAs (one per SM), each walking a CONTIGUOUS range of 128-pixel output tiles; the layer's weights are
//     loaded into shared memory ONCE per CTA and stay resident (layers whose weight image exceeds shared memory split their
//     output channels over CTAs: NSPLIT).  Three pipeline stages run concurrently on different tiles:
//        7 producer warps   stage tile i+1: raw fp32 input -> BN / ReLU -> hi / lo split -> STS.128 into a ring of NBUF (tile, channel
//                           phase) buffers.  The raw input comes either through registers (LDG.256 issued one unit AHEAD, software
//                           pipelined) or, for the layers with shared memory to spare (T2Cfg::RAW), from a ring of RAW raw boxes that
//                           the MMA warp's elected lane fills by TMA tensor loads RAW buffers ahead,
//        1 MMA warp         multiplies tile i (one elected lane; accumulator = TMEM stage i & 1),
//        4 epilogue warps   drain tile i-1: tcgen05.ld -> scale -> 128B-swizzled staging tile in shared memory ->
//                           ONE TMA tensor store per 32-channel half (cp.async.bulk.tensor, clipped at the image border by
//                           the hardware) + per-channel statistics read back column-wise from the staging tile.
//   * No im2col, as before: the halo tile is staged as channel-chunk planes [cin/8][pixel][8 halfs], which IS the canonical
//     K-major no-swizzle UMMA layout, so filter tap (ky, kx) is the same buffer with the descriptor start address advanced
//     by (ky * tile_width + kx) * 16 B; stride-2 layers stage four parity planes.  1x1 layers tile the frame linearly.
//   * Train-mode BatchNorm statistics: per-tile partial sums -> global scratch with plain stores; a CTA publishes them ONCE per
//     frame it touched (its tile range is contiguous: one or two frames) with a barrier + one fence + one ticket atomic; the
//     CTA that completes a frame folds the frame's partials in fixed order (double).
//   * Programmatic dependent launch: the kernel is launched with programmatic stream serialization; everything that touches the
//     previous kernel's data sits behind griddepcontrol.wait, so the prologue overlaps the previous layer's tail.
#include <cuda.h>
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
constexpr int T2_PROD_WARPS = 7;             // 12 warps in all: the register file is allocated per 4 warps, 13 warps would cap a thread at 128 registers
constexpr int T2_PROD = 32 * T2_PROD_WARPS;  // producer threads
constexpr int T2_EPI = 128;                  // epilogue threads (warps 0-3: TMEM lane quadrant = warp id)
constexpr int T2_THREADS = T2_EPI + 32 + T2_PROD;   // warps 0-3 epilogue, warp 4 MMA issuer, warps 5-11 producers

// CIN: real input channels; CINP: padded to a multiple of 16 (zero chunks); COUT: real outputs; NSPLIT: output-channel
// groups processed by different CTAs; CSTAGE: input channels per staged buffer (channel phase); NBUF: ring depth;
// UNR: 32-byte global loads in flight per producer thread and unit.
template <int CIN_, int COUT_, int KS_, int S_, int CSTAGE_, int NBUF_, int NSPLIT_, int UNR_, int RAW_ = 0>
struct T2Cfg {
  static constexpr int CIN = CIN_, COUT = COUT_, KS = KS_, S = S_, CSTAGE = CSTAGE_, NBUF = NBUF_, NSPLIT = NSPLIT_, UNR = UNR_, RAW = RAW_;
  static constexpr int CINP = (CIN + 15) / 16 * 16;
  static constexpr int NOUT = COUT / NSPLIT;                                      // real output channels per CTA
  static constexpr int NP = NOUT <= 32 ? 32 : (NOUT + 15) / 16 * 16;              // padded outputs; UMMA N = 2 * NP (hi rows | lo rows) and NP
  static constexpr int ACC_STRIDE = 2 * NP <= 64 ? 64 : (2 * NP <= 128 ? 128 : 256);   // TMEM columns per accumulator stage
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
  static constexpr int W_UNIT_BYTES = CSTAGE * 2 * NP * 2;                        // [hi rows | lo rows] of one (phase, tap)
  static constexpr int W_BYTES = NPHASE * TAPS * W_UNIT_BYTES;                    // resident weight image of one output group
  static constexpr uint32_t LBO_A = PS, SBO_A = WT * 16;
  static constexpr uint32_t LBO_B = (2 * NP / 8) * 128, SBO_B = 128;
  // A single-phase layer with padded input channels (block2: 24 -> 32) zeroes the padding chunk of every ring buffer ONCE: only the
  // KCR real chunks are staged per tile, by the first PT producer threads (a multiple of KCR, so a thread keeps its chunk).
  static constexpr int KCR = (NPHASE == 1 && CINP != CIN) ? CIN / 8 : KCS;
  static constexpr int PT = T2_PROD - T2_PROD % KCR;
  static constexpr int ITEMS = NSUB * NPIX * KCR;                                 // (pixel, 8-channel chunk) items per staged buffer
  static constexpr int ROUNDS = (ITEMS + PT * UNR - 1) / (PT * UNR);              // units per staged buffer
  static constexpr int NSLOT = ROUNDS * UNR;                                      // items per thread and staged buffer
  static constexpr bool HAS_STG = COUT != 65;                                     // keypoint_head.3 writes its own folded output
  static constexpr int NHALF = HAS_STG ? (NOUT + 31) / 32 : 0;                    // 32-channel TMA boxes per tile
  static constexpr int STG_BYTES = NHALF * 16384;                                 // staging tile(s): [128 px][32 ch] fp32, 128B-swizzled
  static constexpr int FOLD_BYTES = 128 * 2 * 8;
  static constexpr int BN_BYTES = CINP * 2 * 4;                                   // the current frame's (16 rstd, -16 mean rstd) per input channel
  // RAW > 0: the raw fp32 input of a (tile, channel phase) arrives by ONE TMA tensor load (cp.async.bulk.tensor, zero-filled outside the
  // image and beyond CIN) into a ring of RAW shared-memory stages, issued RAW buffers ahead of the tensor core -- the bytes in flight
  // no longer depend on producer registers (one unit ahead = 3 MB in flight chip-wide = 1 TB/s at 2.5 us latency, measured on block2).
  static constexpr int CB = CIN < CSTAGE ? CIN : CSTAGE;                          // channels per raw box
  static constexpr int RAW_W = KS == 1 ? 128 : (S == 1 ? WT : 2 * WT);
  static constexpr int RAW_H = KS == 1 ? 1 : (S == 1 ? HT : 2 * HT);
  static constexpr int RAW_TX = RAW_W * RAW_H * CB * 4;                            // bytes one load delivers
  static constexpr int RAW_BYTES = RAW > 0 ? (RAW_TX + 127) / 128 * 128 : 0;
  static constexpr size_t SMEM_BYTES = (size_t)STG_BYTES + W_BYTES + (size_t)RAW * RAW_BYTES + (size_t)NBUF * BUF_BYTES + FOLD_BYTES + BN_BYTES + 320;
  // kind::f16, F16 x F16 -> F32, both K-major, M = 128
  static constexpr uint32_t IDESC_2N = (1u << 4) | ((uint32_t)(2 * NP >> 3) << 17) | ((128u >> 4) << 24);
  static constexpr uint32_t IDESC_1N = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((128u >> 4) << 24);
  static_assert(CSTAGE % 16 == 0 && CINP % CSTAGE == 0 && COUT % NSPLIT == 0 && (NOUT % 8 == 0 || NOUT == 65) && NOUT <= 80, "shape");
  static_assert(2 * NP <= 256 && (2 * NP) % 16 == 0, "UMMA N");
  static_assert(S == 1 || KS == 3, "stride 2 is implemented for 3x3 only");
  static_assert(T2_PROD % KCS == 0 && PT % KCR == 0 && ROUNDS <= 2, "a producer thread keeps one channel chunk; at most two units per buffer");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(RAW <= 8 && CB % 8 == 0 && RAW_W <= 256 && RAW_H <= 256, "raw input ring");
  static_assert(PS % 16 == 0 && W_UNIT_BYTES % 16 == 0 && W_BYTES % 1024 == 0, "descriptor alignment");
};

struct ConvTc2Args {
  const float* in;          // NHWC raw producer output (or plain values), or xn for T2IN_UNFOLD
  const unsigned char* wimg;   // [split][phase][tap][cstage/8][2*np/8][8][8 halfs]  (rows 0..np-1 = hi, np..2np-1 = lo)
  const float* bias;
  float* out;               // keypoint_head.3 only (every other layer stores through the tensor map)
  const float* in_mean; const float* in_rstd;
  const float* skip_avg; const float* skip_w; const float* skip_b;
  float* part;              // per-tile channel sums [B][tiles][COUT][2] (float)
  unsigned int* ticket; float* out_mean; float* out_rstd;
  int B, Hin, Win, Hout, Wout;
  int full_w;               // T2IN_UNFOLD / KPSOFTMAX: width of xn / K1h
  int tiles_x, tiles;       // tiles per frame
  float out_scale;          // 1 / (activation scale * weight scale), an exact power of two
  unsigned long long* dbg;  // XFB_T2_DEBUG: cycle counters of CTA 0 (32 per layer), else nullptr
  int pdl;                  // launched with programmatic stream serialization (see the kernel)
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

// TMA tensor stores (shared -> global through a CUtensorMap; out-of-range parts of the box are clipped by the hardware)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1),
               "r"(c2), "r"(smem)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3), "r"(smem)
               : "memory");
}
// TMA tensor loads (global -> shared, completion on an mbarrier; parts of the box outside the tensor arrive as zeros)
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t smem, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t smem, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One unit of producer work held in registers: UNR (pixel, 8-channel chunk) items of one staged buffer.
template <int UNR>
struct ProdRegs {
  f8 v[UNR];
  float av[UNR];
  int off[UNR];           // byte offset inside the hi plane set, -1 = no item
  unsigned inside;        // bit u: the item lies inside the image (else zeros are staged)
};
struct ProdPos { int tile, ph, rnd, it; };   // it = staged-buffer counter of this CTA

template <class C, int INMODE, int OUTMODE>
__global__ void __launch_bounds__(T2_THREADS, 1) conv_tc2_kernel(const ConvTc2Args a, const __grid_constant__ CUtensorMap tmap,
                                                                 const __grid_constant__ CUtensorMap tmap_in) {
  constexpr bool USE_RAW = C::RAW > 0 && INMODE != T2IN_UNFOLD;
  constexpr int KCS = C::KCS, PS = C::PS, WT = C::WT, NPIX = C::NPIX, TAPS = C::TAPS, NBUF = C::NBUF, NPHASE = C::NPHASE, UNR = C::UNR,
                ROUNDS = C::ROUNDS, ITEMS = C::ITEMS;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sStg = smem_raw;                                 // staging tile(s) for the TMA store (1024-byte aligned: 128B swizzle)
  unsigned char* sW = smem_raw + C::STG_BYTES;                    // resident weights of this CTA's output group
  unsigned char* sRaw = sW + C::W_BYTES;                          // RAW x raw fp32 input boxes (TMA destinations, 128-byte aligned)
  unsigned char* sA = sRaw + (size_t)C::RAW * C::RAW_BYTES;       // NBUF x [hi | lo] staged tiles
  double* sFold = reinterpret_cast<double*>(sA + (size_t)NBUF * C::BUF_BYTES);   // [128][2]
  float* sBN = reinterpret_cast<float*>(sFold + 128 * 2);                        // [2][CINP]: scale, shift of the producers' current frame
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBN + 2 * C::CINP);
  uint64_t* bar_in = bars + 0;          // [NBUF] tile staged (T2_PROD arrivals)
  uint64_t* bar_free = bars + 4;        // [NBUF] tensor core finished reading the staged buffer
  uint64_t* bar_accf = bars + 8;        // [2] accumulator complete
  uint64_t* bar_acce = bars + 10;       // [2] accumulator drained (4 arrivals)
  uint64_t* bar_w = bars + 12;          // [NPHASE <= 8] weights of a channel phase landed
  uint64_t* bar_raw = bars + 20;        // [RAW <= 8] raw input box landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 28);
  uint32_t* s_flag = s_tmem + 1;

  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const long long t_kernel0 = clock64();
  // work items: (tile of a frame, output group).  A CTA keeps ONE output group (its weights stay resident) and walks a
  // contiguous range of the tiles of all frames (neighbouring tiles share halo rows in L2 and BatchNorm parameters in L1).
  const int split = blockIdx.x % C::NSPLIT;
  const int gidx = blockIdx.x / C::NSPLIT, gcnt = gridDim.x / C::NSPLIT;
  const int n_tiles = a.B * a.tiles;
  const int tile_begin = (int)(((long long)gidx * n_tiles) / gcnt), tile_end = (int)(((long long)(gidx + 1) * n_tiles) / gcnt);

  if (t == 0) {
    for (int s = 0; s < NBUF; ++s) { mbar_init(bar_in + s, T2_PROD); mbar_init(bar_free + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_accf + s, 1); mbar_init(bar_acce + s, 4); }
    for (int s = 0; s < NPHASE; ++s) mbar_init(bar_w + s, 1);
    for (int s = 0; s < C::RAW; ++s) mbar_init(bar_raw + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(s_tmem, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // Programmatic dependent launch: the NEXT layer's CTAs may take an SM as soon as this CTA leaves it and run their prologue
  // (barriers, TMEM, slot tables, weight image load) while other SMs still finish this layer; everything that reads or writes data
  // of the previous kernel sits behind pdl_wait() (= that grid has completed and its writes are visible).
  if (a.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  auto pdl_wait = [&]() { if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory"); };

  if (warp >= 5) {
    // ===================== producers: stage (tile, channel phase) buffers, global loads one unit ahead =====================
    const int pt = t - (T2_EPI + 32);
    const int kc = pt % C::KCR;                      // this thread's 8-channel chunk inside a phase (constant)
    // Per-thread item table, computed ONCE: the (pixel, chunk) slots a thread stages are the same for every tile -- only the tile
    // origin moves.  slot_off = byte offset in the staged buffer (-1: no item), slot_yx = input pixel relative to the tile origin
    // ((dy << 16) | (dx & 0xffff); 1x1 layers: the pixel index inside the 128-pixel tile).
    int slot_off[C::NSLOT], slot_yx[C::NSLOT];
    int slot_raw[USE_RAW ? C::NSLOT : 1];            // byte offset of the item's 8 channels inside the raw box
#pragma unroll
    for (int sl = 0; sl < C::NSLOT; ++sl) {
      const int idx = sl * C::PT + pt;
      slot_off[sl] = -1; slot_yx[sl] = 0;
      if (pt < C::PT && idx < ITEMS) {
        const int rest = idx / C::KCR;               // idx % KCR == kc
        const int pix = rest % NPIX, sub = rest / NPIX;
        slot_off[sl] = sub * C::SUB_BYTES + kc * PS + pix * 16;
        int dy, dx;
        if (C::KS == 1) { dy = 0; dx = pix; }
        else if (C::S == 1) { dy = pix / WT - C::PAD; dx = pix % WT - C::PAD; }
        else { dy = 2 * (pix / WT - 1) + (sub >> 1); dx = 2 * (pix % WT - 1) + (sub & 1); }
        slot_yx[sl] = (int)(((unsigned int)dy << 16) | ((unsigned int)dx & 0xffffu));
        if (USE_RAW) {
          const int brow = (C::KS == 1) ? 0 : (C::S == 1 ? dy + C::PAD : dy + 2), bcol = (C::KS == 1) ? pix : (C::S == 1 ? dx + C::PAD : dx + 2);
          slot_raw[sl] = ((brow * C::RAW_W + bcol) * C::CB + kc * 8) * 4;
        }
      } else if (USE_RAW) slot_raw[sl] = 0;
    }
    if (C::KCR != KCS) {
      // zero the padding chunk planes of every ring buffer once (hi and lo)
      for (int i = pt; i < NBUF * 2 * C::NSUB * (KCS - C::KCR) * (PS / 16); i += T2_PROD) {
        const int per = (KCS - C::KCR) * (PS / 16);
        const int plane = i / per, w = i - plane * per;        // plane = (buffer, hi|lo, sub)
        *reinterpret_cast<uint4*>(sA + (size_t)plane * C::SUB_BYTES + (size_t)C::KCR * PS + (size_t)w * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    const bool dbg_me = a.dbg != nullptr && blockIdx.x == 0 && pt == 0;
    long long t0 = dbg_me ? clock64() : 0;

    // L2 prefetch of the input rows a tile reads (issued two tiles ahead by the first threads: one row each)
    auto prefetch_tile = [&](int tile) {
      const int b = tile / a.tiles, tf = tile - b * a.tiles;
      if (INMODE == T2IN_UNFOLD) {
        // 128 consecutive cells -> the xn rows of the cell rows they touch (full rows: full_w floats)
        const int cy0 = (tf * 128) / a.Win, cy1 = min(a.Hin - 1, (tf * 128 + 127) / a.Win);
        const int nrow = (cy1 - cy0 + 1) * 8;
        if (pt < nrow) {
          const float* p = a.in + ((size_t)b * (a.Hin * 8) + (size_t)cy0 * 8 + pt) * a.full_w;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)a.full_w * 4u) : "memory");
        }
      } else if (C::KS == 1) {
        const int px0 = tf * 128, npx = min(128, a.Hin * a.Win - px0);
        if (pt < 8 && npx > 0) {      // 8 slices of the contiguous pixel range
          const int per = (npx + 7) / 8, s0 = pt * per, cnt = min(per, npx - s0);
          if (cnt > 0) {
            const float* p = a.in + ((size_t)b * a.Hin * a.Win + px0 + s0) * C::CIN;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)(cnt * C::CIN * 4)) : "memory");
          }
        }
      } else {
        const int oy0 = (tf / a.tiles_x) * C::TH, ox0 = (tf % a.tiles_x) * C::TW;
        const int y0 = (C::S == 1) ? oy0 - 1 : 2 * oy0 - 1, ny = (C::S == 1) ? C::TH + 2 : 2 * C::TH + 1;
        const int x0 = max(0, (C::S == 1) ? ox0 - 1 : 2 * ox0 - 1), x1 = min(a.Win, (C::S == 1) ? ox0 + C::TW + 1 : 2 * (ox0 + C::TW));
        const int iy = y0 + pt;
        if (pt < ny && iy >= 0 && iy < a.Hin && x1 > x0) {
          const float* p = a.in + (((size_t)b * a.Hin + iy) * a.Win + x0) * C::CIN;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)((x1 - x0) * C::CIN * 4)) : "memory");
        }
      }
    };
    // decode a unit and issue its global loads
    auto issue = [&](const ProdPos& u, ProdRegs<UNR>& R) {
      const int b = u.tile / a.tiles, tf = u.tile - b * a.tiles;
      const int oy0 = (C::KS == 1) ? 0 : (tf / a.tiles_x) * C::TH, ox0 = (C::KS == 1) ? 0 : (tf % a.tiles_x) * C::TW;
      const float* in_b = (INMODE == T2IN_UNFOLD) ? a.in + (size_t)b * (a.Hin * 8) * a.full_w : a.in + (size_t)b * a.Hin * a.Win * C::CIN;
      const int ch = u.ph * C::CSTAGE + kc * 8;      // first of this thread's 8 channels
      const bool ch_ok = ch < C::CIN;                // padded chunks (Cin 24 -> 32) are zero
      R.inside = 0u;
      int uy0 = 0, ux0 = 0;                          // T2IN_UNFOLD: cell of the tile's first pixel
      if (INMODE == T2IN_UNFOLD) { uy0 = (tf * 128) / a.Win; ux0 = tf * 128 - uy0 * a.Win; }
      const int lin_ok = (C::KS == 1) ? a.Hin * a.Win - tf * 128 : 0;          // 1x1: pixels of the frame left from this tile on
      const int oyi = (C::S == 1) ? oy0 : 2 * oy0, oxi = (C::S == 1) ? ox0 : 2 * ox0;
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        const int off = (ROUNDS == 1 || u.rnd == 0) ? slot_off[q] : slot_off[(ROUNDS - 1) * UNR + q];
        const int yx = (ROUNDS == 1 || u.rnd == 0) ? slot_yx[q] : slot_yx[(ROUNDS - 1) * UNR + q];
        R.v[q].a = make_float4(0.f, 0.f, 0.f, 0.f); R.v[q].b = R.v[q].a;
        R.av[q] = 0.f; R.off[q] = off;
        if (off >= 0 && ch_ok) {
          const int dy = yx >> 16, dx = (int)(short)(yx & 0xffff);
          if (C::KS == 1 && INMODE != T2IN_UNFOLD) {
            if (dx < lin_ok) {                       // the tile is 128 consecutive pixels of the frame
              R.inside |= 1u << q;
              if (!USE_RAW) R.v[q] = ldg256(in_b + ((size_t)tf * 128 + dx) * C::CIN + ch);
            }
          } else {
            int iy, ix;
            if (INMODE == T2IN_UNFOLD) {
              ix = ux0 + dx; iy = uy0;
              if (a.Win >= 128) { if (ix >= a.Win) { ix -= a.Win; ++iy; } }
              else { const int wr = ix / a.Win; iy += wr; ix -= wr * a.Win; }
            } else { iy = oyi + dy; ix = oxi + dx; }
            if (iy >= 0 && iy < a.Hin && ix >= 0 && ix < a.Win) {
              R.inside |= 1u << q;
              if (INMODE == T2IN_UNFOLD) {
                // XFeatModel::unfold2d(x, 8), src/XFeat.cc:124-133: channel c = (y % 8) * 8 + x % 8 -> chunk kc = row iy * 8 + kc of xn
                R.v[q] = ldg256(in_b + (size_t)(iy * 8 + (ch >> 3)) * a.full_w + ix * 8);
              } else {
                if (!USE_RAW) R.v[q] = ldg256(in_b + ((size_t)iy * a.Win + ix) * C::CIN + ch);
                if (INMODE == T2IN_BN_SKIP) R.av[q] = a.skip_avg[((size_t)b * a.Hin + iy) * a.Win + ix];
              }
            }
          }
        }
      }
    };
    // BatchNorm parameters of a unit's 8 channels, from shared memory.  The producers keep the CURRENT frame's parameters of all
    // input channels there in fused form (scale = 16 rstd, shift = -16 mean rstd: relu((x - mean) rstd) * 16 = max(fma(x, scale,
    // shift), 0), the scale / shift form ATen's batch_norm uses, with the exact activation scale folded in) and reload them -- two
    // named barriers among the 224 producer threads -- only when their tile range enters another frame: no global load, no
    // dependent latency per staged buffer.
    float m[8], r[8], sw[8], sb[8];
    int cur_b = -1, sw_loaded = 0;
    auto load_params = [&](const ProdPos& u) {
      if (!(INMODE == T2IN_BN || INMODE == T2IN_BN_SKIP)) return;
      const int b = u.tile / a.tiles;
      if (C::KCR != KCS) {
        // block2 (single phase, 24 channels): measured faster with the parameters straight from global memory / L1, loaded BEFORE the
        // next unit's data loads are issued (the L1 returns a warp's loads in order: behind them they would wait for DRAM latency)
        const int ch = kc * 8;
        const f8 mm = ldg256(a.in_mean + b * C::CIN + ch), rr = ldg256(a.in_rstd + b * C::CIN + ch);
        m[0] = mm.a.x; m[1] = mm.a.y; m[2] = mm.a.z; m[3] = mm.a.w; m[4] = mm.b.x; m[5] = mm.b.y; m[6] = mm.b.z; m[7] = mm.b.w;
        r[0] = rr.a.x; r[1] = rr.a.y; r[2] = rr.a.z; r[3] = rr.a.w; r[4] = rr.b.x; r[5] = rr.b.y; r[6] = rr.b.z; r[7] = rr.b.w;
#pragma unroll
        for (int e = 0; e < 8; ++e) { r[e] *= T2_ACT_SCALE; m[e] = -m[e] * r[e]; }
        if (INMODE == T2IN_BN_SKIP && sw_loaded == 0) {
          const f8 ww = ldg256(a.skip_w + ch), bb = ldg256(a.skip_b + ch);
          sw[0] = ww.a.x; sw[1] = ww.a.y; sw[2] = ww.a.z; sw[3] = ww.a.w; sw[4] = ww.b.x; sw[5] = ww.b.y; sw[6] = ww.b.z; sw[7] = ww.b.w;
          sb[0] = bb.a.x; sb[1] = bb.a.y; sb[2] = bb.a.z; sb[3] = bb.a.w; sb[4] = bb.b.x; sb[5] = bb.b.y; sb[6] = bb.b.z; sb[7] = bb.b.w;
          sw_loaded = 1;
        }
        return;
      }
      if (b != cur_b) {                             // (uniform over the producer threads: they walk the same unit sequence)
        asm volatile("bar.sync 2, %0;" ::"n"(T2_PROD) : "memory");      // everybody is done reading the previous frame's parameters
        for (int c = pt; c < C::CINP; c += T2_PROD) {
          float sc = 0.f, sh = 0.f;
          if (c < C::CIN) { sc = a.in_rstd[b * C::CIN + c] * T2_ACT_SCALE; sh = -a.in_mean[b * C::CIN + c] * sc; }
          sBN[c] = sc; sBN[C::CINP + c] = sh;
        }
        asm volatile("bar.sync 2, %0;" ::"n"(T2_PROD) : "memory");
        cur_b = b;
      }
      const int ch = u.ph * C::CSTAGE + kc * 8;
      const float4 r0 = *reinterpret_cast<const float4*>(sBN + ch), r1 = *reinterpret_cast<const float4*>(sBN + ch + 4);
      const float4 m0 = *reinterpret_cast<const float4*>(sBN + C::CINP + ch), m1 = *reinterpret_cast<const float4*>(sBN + C::CINP + ch + 4);
      r[0] = r0.x; r[1] = r0.y; r[2] = r0.z; r[3] = r0.w; r[4] = r1.x; r[5] = r1.y; r[6] = r1.z; r[7] = r1.w;
      m[0] = m0.x; m[1] = m0.y; m[2] = m0.z; m[3] = m0.w; m[4] = m1.x; m[5] = m1.y; m[6] = m1.z; m[7] = m1.w;
      if (INMODE == T2IN_BN_SKIP && cur_b >= 0 && sw_loaded == 0 && ch < C::CIN) {
        const f8 ww = ldg256(a.skip_w + ch), bb = ldg256(a.skip_b + ch);     // (single-phase layer: the chunk, hence these, never change)
        sw[0] = ww.a.x; sw[1] = ww.a.y; sw[2] = ww.a.z; sw[3] = ww.a.w; sw[4] = ww.b.x; sw[5] = ww.b.y; sw[6] = ww.b.z; sw[7] = ww.b.w;
        sb[0] = bb.a.x; sb[1] = bb.a.y; sb[2] = bb.a.z; sb[3] = bb.a.w; sb[4] = bb.b.x; sb[5] = bb.b.y; sb[6] = bb.b.z; sb[7] = bb.b.w;
        sw_loaded = 1;
      }
    };
    // transform the loaded unit and store it into its staged buffer
    auto consume = [&](const ProdPos& u, const ProdRegs<UNR>& R) {
      const int buf = u.it % NBUF;
      unsigned char* dst_hi = sA + (size_t)buf * C::BUF_BYTES;
      const unsigned char* raw = sRaw + (size_t)(USE_RAW ? u.it % (C::RAW > 0 ? C::RAW : 1) : 0) * C::RAW_BYTES;
      if (u.rnd == 0) {
        if (!USE_RAW && u.ph == 0 && u.tile + 2 < tile_end) prefetch_tile(u.tile + 2);
        T2_TICK(3);
        if (u.it >= NBUF) mbar_wait(bar_free + buf, ((u.it / NBUF) - 1) & 1);   // the MMAs that read this buffer are done
        if (USE_RAW) mbar_wait(bar_raw + u.it % (C::RAW > 0 ? C::RAW : 1), (u.it / (C::RAW > 0 ? C::RAW : 1)) & 1);   // the raw box has landed
        T2_TICK(2);
      }
#pragma unroll
      for (int q = 0; q < UNR; ++q) {
        if (R.off[q] < 0) continue;
        float x[8] = {R.v[q].a.x, R.v[q].a.y, R.v[q].a.z, R.v[q].a.w, R.v[q].b.x, R.v[q].b.y, R.v[q].b.z, R.v[q].b.w};
        if (USE_RAW) {
          const int ro = (ROUNDS == 1 || u.rnd == 0) ? slot_raw[USE_RAW ? q : 0] : slot_raw[USE_RAW ? (ROUNDS - 1) * UNR + q : 0];
          float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
          if ((R.inside >> q) & 1u) { x0 = *reinterpret_cast<const float4*>(raw + ro); x1 = *reinterpret_cast<const float4*>(raw + ro + 16); }
          x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
        }
        if ((R.inside >> q) & 1u) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (INMODE == T2IN_BN || INMODE == T2IN_BN_SKIP) {
              x[e] = fmaxf(fmaf(x[e], r[e], m[e]), 0.f);
              // x1 + skip1(x), src/XFeat.cc:153; skip1 = AvgPool2d(4,4) + Conv2d(1,24,1) (:36-39)
              if (INMODE == T2IN_BN_SKIP) x[e] = fmaf(fmaf(R.av[q], sw[e], sb[e]), T2_ACT_SCALE, x[e]);
            } else {
              x[e] *= T2_ACT_SCALE;
            }
          }
        }
        uint4 hi, lo;
        split2(x[0], x[1], hi.x, lo.x); split2(x[2], x[3], hi.y, lo.y); split2(x[4], x[5], hi.z, lo.z); split2(x[6], x[7], hi.w, lo.w);
        *reinterpret_cast<uint4*>(dst_hi + R.off[q]) = hi;
        *reinterpret_cast<uint4*>(dst_hi + C::IN_BYTES + R.off[q]) = lo;
      }
      if (u.rnd == ROUNDS - 1) {
        T2_TICK(3);
        fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
        mbar_arrive(bar_in + buf);
        T2_TICK(4);
        if (dbg_me) atomicAdd(a.dbg + 5, 1ull);
      }
    };
    auto advance = [&](ProdPos& u) {
      if (++u.rnd == ROUNDS) {
        u.rnd = 0; ++u.it;
        if (++u.ph == NPHASE) { u.ph = 0; ++u.tile; }
      }
    };

    ProdRegs<UNR> ra, rb;
    ProdPos u = {tile_begin, 0, 0, 0};
    pdl_wait();
    if (u.tile < tile_end) {
      issue(u, ra);
      while (true) {
        ProdPos un = u;
        advance(un);
        bool more = un.tile < tile_end;
        load_params(u);
        if (more) issue(un, rb);           // the next unit's loads fly while this one is transformed
        consume(u, ra);
        if (!more) break;
        u = un;
        advance(un);
        more = un.tile < tile_end;
        load_params(u);
        if (more) issue(un, ra);
        consume(u, rb);
        if (!more) break;
        u = un;
      }
    }
  } else if (warp == 4) {
    // ===================== weights (once) + MMA issue: the whole warp runs the uniform loop, ONE elected lane issues ===========
    const bool leader = elect_one_sync();
    // raw input boxes: buffer it2 = (tile, channel phase) of this CTA's walk goes to ring stage it2 % RAW
    const int total_its = (tile_end - tile_begin) * NPHASE;
    auto raw_load = [&](int it2) {
      const int tile = tile_begin + it2 / NPHASE, ph = it2 % NPHASE;
      const int b = tile / a.tiles, tf = tile - b * a.tiles;
      const int stage = it2 % (C::RAW > 0 ? C::RAW : 1);
      mbar_expect_tx(bar_raw + stage, C::RAW_TX);
      const uint32_t dst = smem_u32(sRaw + (size_t)stage * C::RAW_BYTES);
      if (C::KS == 1) tma_load_3d(&tmap_in, dst, bar_raw + stage, ph * C::CSTAGE, tf * 128, b);
      else {
        const int oy0 = (tf / a.tiles_x) * C::TH, ox0 = (tf % a.tiles_x) * C::TW;
        if (C::S == 1) tma_load_4d(&tmap_in, dst, bar_raw + stage, ph * C::CSTAGE, ox0 - C::PAD, oy0 - C::PAD, b);
        else tma_load_4d(&tmap_in, dst, bar_raw + stage, ph * C::CSTAGE, 2 * ox0 - 2, 2 * oy0 - 2, b);
      }
    };
    if (leader && tile_begin < tile_end) {
      const unsigned char* wsrc = a.wimg + (size_t)split * C::W_BYTES;        // (weights do not depend on the previous kernel)
      constexpr uint32_t PH_BYTES = (uint32_t)TAPS * C::W_UNIT_BYTES;
      for (int ph = 0; ph < NPHASE; ++ph) {
        mbar_expect_tx(bar_w + ph, PH_BYTES);
        for (int u = 0; u < TAPS; ++u)
          bulk_g2s(sW + (size_t)ph * PH_BYTES + (size_t)u * C::W_UNIT_BYTES, wsrc + (size_t)ph * PH_BYTES + (size_t)u * C::W_UNIT_BYTES, C::W_UNIT_BYTES,
                   bar_w + ph);
      }
    }
    pdl_wait();
    if (USE_RAW && leader)
      for (int i = 0; i < C::RAW && i < total_its; ++i) raw_load(i);
    __syncwarp();
    const uint64_t da0 = umma_desc_kmajor(smem_u32(sA), C::LBO_A, C::SBO_A);
    const uint64_t db0 = umma_desc_kmajor(smem_u32(sW), C::LBO_B, C::SBO_B);
    constexpr uint64_t KA = (2u * C::LBO_A) >> 4, KB = (2u * C::LBO_B) >> 4;      // one K = 16 step (two 16-byte chunks)
    constexpr uint64_t A_LO = (uint64_t)C::IN_BYTES >> 4, A_BUF = (uint64_t)C::BUF_BYTES >> 4;
    constexpr uint64_t W_UNIT = (uint64_t)C::W_UNIT_BYTES >> 4;
    int it = 0, n = 0;
    const bool dbg_me = a.dbg != nullptr && blockIdx.x == 0 && leader;
    long long t0 = dbg_me ? clock64() : 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++n) {
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
        // every producer has read raw stage it % RAW (they arrived on bar_in after their last read): refill it, RAW buffers ahead
        if (USE_RAW && leader && it + C::RAW < total_its) raw_load(it + C::RAW);
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
            const uint64_t bw = db0 + (uint64_t)(ph * TAPS + tap) * W_UNIT;
#pragma unroll
            for (int k16 = 0; k16 < C::CSTAGE / 16; ++k16) {
              // columns [0, NP) += A_hi W_hi, [NP, 2NP) += A_hi W_lo: one fetch of A_hi for two products; then [0, NP) += A_lo W_hi
              umma_f16_ss(d, ah + k16 * KA, bw + k16 * KB, C::IDESC_2N, (ph > 0 || tap > 0 || k16 > 0) ? 1u : 0u);
              umma_f16_ss(d, al + k16 * KA, bw + k16 * KB, C::IDESC_1N, 1u);
            }
          }
          umma_commit(bar_free + buf);                // the staged buffer may be refilled once these MMAs have read it
          if (ph == NPHASE - 1) umma_commit(bar_accf + acc);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> staging tile -> TMA store (+ statistics) =====================
    const int quad = warp;                              // TMEM lane quadrant (hardware: warp id % 4)
    const int p = quad * 32 + lane;                     // pixel of the tile = TMEM lane
    const int cbase = split * C::NOUT;                  // first output channel of this CTA's group
    int n = 0;
    const bool dbg_me = a.dbg != nullptr && blockIdx.x == 0 && t == 0;
    long long t0 = dbg_me ? clock64() : 0;
    pdl_wait();                                         // (the partial-sum scratch and the ticket counters are shared between layers)
    int run_b = -1, run_cnt = 0;                        // tiles of frame run_b this CTA has written partials for, not yet published
    const int items_per_frame = a.tiles * C::NSPLIT;

    // the frame's last work item: fixed-order fold of all tile partials (slice-strided, then slice order), in double
    auto fold = [&](int b) {
      constexpr int NSL = C::COUT >= 128 ? 1 : (C::COUT >= 64 ? 2 : 4);
      const float* part_b = a.part + (size_t)b * a.tiles * C::COUT * 2;
      const int entries = a.tiles;
      if (t < NSL * C::COUT) {
        const int c = t % C::COUT, sl = t / C::COUT;
        double d1 = 0.0, d2 = 0.0;
        int i = sl;
        for (; i + 7 * NSL < entries; i += 8 * NSL) {          // 8 independent loads in flight, summed in index order
          float2 x[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) x[q] = __ldcg(reinterpret_cast<const float2*>(part_b + ((size_t)(i + q * NSL) * C::COUT + c) * 2));
#pragma unroll
          for (int q = 0; q < 8; ++q) { d1 += (double)x[q].x; d2 += (double)x[q].y; }
        }
        for (; i < entries; i += NSL) {
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
      asm volatile("bar.sync 1, 128;" ::: "memory");     // sFold is reused
    };
    // Publish this CTA's partials of frame b (cnt work items): the barrier orders every thread's partial stores before thread 0's
    // fence + ticket (cumulative at gpu scope); whoever completes the frame folds it.  Once per (CTA, frame), not per tile.
    auto publish = [&](int b, int cnt) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (t == 0) {
        __threadfence();
        const unsigned int prev = atomicAdd(a.ticket + b * XFB_TICKET_STRIDE, (unsigned int)cnt);
        *s_flag = (prev + (unsigned int)cnt == (unsigned int)items_per_frame) ? 1u : 0u;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (*s_flag) { __threadfence(); fold(b); }
    };

    for (int tile = tile_begin; tile < tile_end; ++tile, ++n) {
      const int acc = n & 1;
      const int b = tile / a.tiles, tf = tile - b * a.tiles;
      if (dbg_me) atomicAdd(a.dbg + 16, 1ull);
      if constexpr (OUTMODE == T2OUT_STATS) {
        if (run_b >= 0 && b != run_b) { publish(run_b, run_cnt); run_cnt = 0; }
        run_b = b;
      }
      const int oy0 = (C::KS == 1) ? 0 : (tf / a.tiles_x) * C::TH, ox0 = (C::KS == 1) ? 0 : (tf % a.tiles_x) * C::TW;
      const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * C::ACC_STRIDE;
      T2_TICK(15);
      mbar_wait(bar_accf + acc, (n >> 1) & 1);
      tc_fence_after();
      T2_TICK(10);
      if constexpr (OUTMODE == T2OUT_KPSOFTMAX) {
        // keypoint_head.3 epilogue: + bias, softmax over the 65 logits, drop the dustbin, 8x8 fold
        // (src/XFeat.cc:85-90, XFextractor::getKptsHeatmap src/XFextractor.cc:204-217); one thread per cell
        const int lin = tf * 128 + p;
        const int oy = lin / a.Wout, ox = lin - oy * a.Wout;
        float v[80];
        {
          float w[32];
          tmem_ld32_nw(tq, v); tmem_ld32_nw(tq + (uint32_t)C::NP, w);                 // [0, NP): A_hi W_hi + A_lo W_hi; [NP, 2NP): A_hi W_lo
          tmem_ld_fence();
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] += w[c];
          tmem_ld32_nw(tq + 32u, v + 32); tmem_ld32_nw(tq + (uint32_t)C::NP + 32u, w);
          tmem_ld_fence();
#pragma unroll
          for (int c = 0; c < 32; ++c) v[32 + c] += w[c];
          tmem_ld16(tq + 64u, v + 64); tmem_ld16(tq + (uint32_t)C::NP + 64u, w);
          tmem_ld_fence();
#pragma unroll
          for (int c = 0; c < 16; ++c) v[64 + c] += w[c];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acce + acc);
        float mx = -3.4e38f;
#pragma unroll
        for (int c = 0; c < 65; ++c) { v[c] = fmaf(v[c], a.out_scale, a.bias[c]); mx = fmaxf(mx, v[c]); }
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < 65; ++c) { v[c] = __expf(v[c] - mx); sum += v[c]; }
        const float inv = 1.0f / sum;
        if (oy < a.Hout) {
          float* dst = a.out + (size_t)b * (a.Hout * 8) * a.full_w + (size_t)(oy * 8) * a.full_w + ox * 8;
#pragma unroll
          for (int ry = 0; ry < 8; ++ry)
            stg256(dst + (size_t)ry * a.full_w, v[ry * 8 + 0] * inv, v[ry * 8 + 1] * inv, v[ry * 8 + 2] * inv, v[ry * 8 + 3] * inv, v[ry * 8 + 4] * inv,
                   v[ry * 8 + 5] * inv, v[ry * 8 + 6] * inv, v[ry * 8 + 7] * inv);
        }
      } else {
        constexpr int NCH = C::NP / 32;                 // 32-column chunks of real outputs (NP is 32 or 64 here)
        static_assert(C::NP % 32 == 0, "epilogue chunking");
        float v[NCH][32];
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          float w[32];
          tmem_ld32_nw(tq + (uint32_t)ci * 32u, v[ci]);
          tmem_ld32_nw(tq + (uint32_t)C::NP + (uint32_t)ci * 32u, w);       // the A_hi x W_lo columns
          tmem_ld_fence();
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            float x = (v[ci][q] + w[q]) * a.out_scale;
            if (OUTMODE == T2OUT_BIAS && ci * 32 + q < C::NOUT) x += a.bias[cbase + ci * 32 + q];
            v[ci][q] = x;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acce + acc);     // values are in registers: the tensor core may reuse the accumulator
        T2_TICK(11);
        // the staging tile is free once the previous tile's TMA store has READ it (and every thread is done with its column sums)
        if (t == 0) bulk_wait_read0();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        T2_TICK(14);
        // registers -> staging: [128 px][32 ch] fp32 per 32-channel half, 16-byte chunk c of pixel row p at chunk (c ^ (p & 7))
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          unsigned char* row = sStg + (size_t)ci * 16384 + (size_t)p * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(row + ((c ^ (p & 7)) << 4)) = make_float4(v[ci][4 * c], v[ci][4 * c + 1], v[ci][4 * c + 2], v[ci][4 * c + 3]);
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (t == 0) {
#pragma unroll
          for (int h = 0; h < C::NHALF; ++h) {
            if (C::KS == 1) tma_store_3d(&tmap, smem_u32(sStg + (size_t)h * 16384), cbase + h * 32, tf * 128, b);
            else tma_store_4d(&tmap, smem_u32(sStg + (size_t)h * 16384), cbase + h * 32, ox0, oy0, b);
          }
          bulk_commit();
        }
        T2_TICK(12);
        if constexpr (OUTMODE == T2OUT_STATS) {
          // per-channel sum / sum of squares over the valid pixels, read back from the staging tile with 16-byte loads: lane =
          // (channel chunk of 4, pixel octet o: the pixels o, o + 8, ...), a warp owns 4 chunks; octets are paired (o, o + 4) inside a
          // quarter warp so that the swizzled chunk positions never collide; 3 shuffle steps merge the 8 octets
          {
            constexpr int NCHUNK = C::NOUT / 4;                         // 16-byte chunks of real channels (6, 8 or 16)
            const int o = ((lane >> 2) & 1) * 4 + (lane >> 3);          // pixel octet of this lane
            int rows_ok = 16, cols_ok = 8, lin_ok = 128;
            if (C::KS == 1) lin_ok = min(128, a.Hout * a.Wout - tf * 128);
            else { rows_ok = min(16, a.Hout - oy0); cols_ok = min(8, a.Wout - ox0); }
            float* part_b = a.part + (size_t)b * a.tiles * C::COUT * 2;
#pragma unroll
            for (int rep = 0; rep < (NCHUNK + 15) / 16; ++rep) {
              const int cc = rep * 16 + quad * 4 + (lane & 3);          // chunk of 4 channels
              float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
              if (cc < NCHUNK) {
                const unsigned char* base = sStg + (size_t)(cc >> 3) * 16384 + (size_t)o * 128 + (((cc & 7) ^ o) << 4);
                float4 x[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = *reinterpret_cast<const float4*>(base + (size_t)i * 1024);   // pixel o + 8 i = tile row i, column o
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  const bool ok = (C::KS == 1) ? (o + 8 * i < lin_ok) : (i < rows_ok && o < cols_ok);
                  const float4 y = ok ? x[i] : make_float4(0.f, 0.f, 0.f, 0.f);                                  // (a zero adds nothing to either sum)
                  s1.x += y.x; s1.y += y.y; s1.z += y.z; s1.w += y.w;
                  s2.x = fmaf(y.x, y.x, s2.x); s2.y = fmaf(y.y, y.y, s2.y); s2.z = fmaf(y.z, y.z, s2.z); s2.w = fmaf(y.w, y.w, s2.w);
                }
              }
#pragma unroll
              for (int off = 4; off < 32; off <<= 1) {                  // lanes differing in bits 2..4 hold the other octets of the same chunk
                s1.x += __shfl_xor_sync(0xffffffffu, s1.x, off); s1.y += __shfl_xor_sync(0xffffffffu, s1.y, off);
                s1.z += __shfl_xor_sync(0xffffffffu, s1.z, off); s1.w += __shfl_xor_sync(0xffffffffu, s1.w, off);
                s2.x += __shfl_xor_sync(0xffffffffu, s2.x, off); s2.y += __shfl_xor_sync(0xffffffffu, s2.y, off);
                s2.z += __shfl_xor_sync(0xffffffffu, s2.z, off); s2.w += __shfl_xor_sync(0xffffffffu, s2.w, off);
              }
              if (lane < 4 && cc < NCHUNK) {
                float* dst = part_b + ((size_t)tf * C::COUT + cbase + cc * 4) * 2;
                *reinterpret_cast<float4*>(dst) = make_float4(s1.x, s2.x, s1.y, s2.y);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(s1.z, s2.z, s1.w, s2.w);
              }
            }
          }
          ++run_cnt;
          T2_TICK(13);
        }
      }
    }
    if constexpr (OUTMODE == T2OUT_STATS) {
      if (run_b >= 0) publish(run_b, run_cnt);
    }
    if (C::HAS_STG && t == 0) bulk_wait0();            // the last stores have left shared memory and are complete
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
//                        CIN COUT KS S CSTAGE NBUF NSPLIT UNR
using T2B2x = T2Cfg<24, 24, 3, 1, 32, 3, 1, 3, 4>;   // block2.0/.1                   120x160         (Cin padded 24 -> 32; raw ring of 4)
using T2B30 = T2Cfg<24, 64, 3, 2, 16, 3, 1, 3>;      // block3.0                      -> 60x80
using T2C33 = T2Cfg<64, 64, 3, 1, 32, 2, 1, 4>;      // block3.1, block4.1/.2, block_fusion.0/.1   (147 KB of weights resident)
using T2C11 = T2Cfg<64, 64, 1, 1, 64, 2, 1, 3, 3>;   // block3.2, block_fusion.2, heatmap_head.0/.1, keypoint_head.0/.1/.2   (raw ring of 3)
using T2B40 = T2Cfg<64, 64, 3, 2, 16, 2, 2, 3>;      // block4.0                      -> 30x40        (two output groups of 32)
using T2B50 = T2Cfg<64, 128, 3, 2, 16, 2, 4, 3>;     // block5.0                      -> 15x20        (four output groups of 32)
using T2B5x = T2Cfg<128, 128, 3, 1, 32, 2, 4, 4>;    // block5.1/.2                                   (four output groups of 32)
using T2B53 = T2Cfg<128, 64, 1, 1, 64, 3, 1, 3>;     // block5.3
using T2KP3 = T2Cfg<64, 65, 1, 1, 64, 3, 1, 3>;      // keypoint_head.3 (65 outputs, N padded to 80) + softmax / fold epilogue

struct T2LayerInfo { int cstage, np, nsplit, cinp, ks; };
template <class C> static T2LayerInfo t2_info_of() { return {C::CSTAGE, C::NP, C::NSPLIT, C::CINP, C::KS}; }
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

// ---- tensor maps of the layer outputs (host) -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static_assert(sizeof(CUtensorMap) == sizeof(Ctx::TmapSlot::blob), "tensor map storage");

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// NHWC fp32 output [B][H][W][C]: 3x3 layers store 8 x 16-pixel tiles (4-D map {C, W, H, B}), 1x1 layers 128 consecutive pixels
// (3-D map {C, H*W, B}); box = 32 channels (128 bytes, 128B swizzle), clipped at the frame border by the hardware.
static cudaError_t output_tmap(Ctx* c, int L, float* out, int B, int H, int W, int C, bool linear, const CUtensorMap** map) {
  Ctx::TmapSlot& s = c->tmap[L];
  if (s.ptr != out || s.B != B || s.H != H || s.W != W) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap m;
    CUresult r;
    if (linear) {
      const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)H * W, (cuuint64_t)B};
      const cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)H * W * C * 4};
      const cuuint32_t box[3] = {32, 128, 1}, es[3] = {1, 1, 1};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
      const cuuint32_t box[4] = {32, 8, 16, 1}, es[4] = {1, 1, 1, 1};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    std::memcpy(s.blob, &m, sizeof(m));
    s.ptr = out; s.B = B; s.H = H; s.W = W;
  }
  *map = reinterpret_cast<const CUtensorMap*>(s.blob);
  return cudaSuccess;
}

// NHWC fp32 input [B][H][W][C] of a layer with a raw ring: box = the halo tile (3x3: {CB, RAW_W, RAW_H, 1} at any, also negative,
// origin) or 128 consecutive pixels (1x1: 3-D map {C, H*W, B}); no swizzle (the producers read 32-byte pieces); what lies outside
// the tensor -- image border, channels beyond C -- arrives as zeros.
static cudaError_t input_tmap(Ctx* c, int L, const float* in, int B, int H, int W, int C, bool linear, int cb, int bw, int bh, const CUtensorMap** map) {
  Ctx::TmapSlot& s = c->tmap_in[L];
  if (s.ptr != in || s.B != B || s.H != H || s.W != W) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap m;
    CUresult r;
    void* base = const_cast<float*>(in);
    if (linear) {
      const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)H * W, (cuuint64_t)B};
      const cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)H * W * C * 4};
      const cuuint32_t box[3] = {(cuuint32_t)cb, (cuuint32_t)bw, 1}, es[3] = {1, 1, 1};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
      const cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)bw, (cuuint32_t)bh, 1}, es[4] = {1, 1, 1, 1};
      r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    std::memcpy(s.blob, &m, sizeof(m));
    s.ptr = in; s.B = B; s.H = H; s.W = W;
  }
  *map = reinterpret_cast<const CUtensorMap*>(s.blob);
  return cudaSuccess;
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
  const CUtensorMap* map = reinterpret_cast<const CUtensorMap*>(c->tmap[tag].blob);     // (unused by keypoint_head.3)
  if (C::HAS_STG) {
    cudaError_t e = output_tmap(c, tag, a.out, a.B, a.Hout, a.Wout, C::COUT, C::KS == 1, &map);
    if (e != cudaSuccess) return e;
  }
  const CUtensorMap* map_in = map;                                                      // (unused without a raw ring)
  if (C::RAW > 0 && INMODE != T2IN_UNFOLD) {
    cudaError_t e = input_tmap(c, tag, a.in, a.B, a.Hin, a.Win, C::CIN, C::KS == 1, C::CB, C::RAW_W, C::RAW_H, &map_in);
    if (e != cudaSuccess) return e;
  }
  const int items = a.B * a.tiles * C::NSPLIT;
  int grid = c->num_sms - c->num_sms % C::NSPLIT;            // persistent: one CTA per SM, a multiple of the output groups
  if (grid > items) grid = items;                            // (items is a multiple of NSPLIT)
  prof_begin(c, tag);
  a.pdl = (c->pdl && !c->prof) ? 1 : 0;                 // (per-kernel event timing wants the launches serialised)
  if (a.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(T2_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, *map, *map_in);
    if (e != cudaSuccess) return e;
  } else {
    kern<<<grid, T2_THREADS, C::SMEM_BYTES, c->stream>>>(a, *map, *map_in);
  }
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

// Host side of xfb_create: OIHW fp32 weights -> [split][phase][tap][cstage/8][2*np/8][8][8 halfs] UMMA operand images of
// w * 2^k (k chosen so that max |w| * 2^k is in [1024, 2048)): rows [0, np) hold the fp16 hi pieces, rows [np, 2np) the lo
// pieces.  Returns the epilogue factor 1 / (16 * 2^k).
float conv_tc2_pack_weights(int L, const float* oihw, int cout, int cin, int ks, std::vector<unsigned char>& img) {
  const T2LayerInfo li = t2_info(L);
  const int taps = ks * ks, nphase = li.cinp / li.cstage, nout = cout / li.nsplit;
  float wmax = 0.f;
  for (size_t i = 0; i < (size_t)cout * cin * taps; ++i) wmax = std::fmax(wmax, std::fabs(oihw[i]));
  int k = 0;
  if (wmax > 0.f) { int e; std::frexp(wmax, &e); k = 11 - e; }           // wmax = f * 2^e, f in [0.5, 1)  ->  wmax * 2^k in [1024, 2048)
  const float wscale = std::ldexp(1.0f, k);
  const size_t unit = (size_t)li.cstage * 2 * li.np * 2;                 // bytes of one (phase, tap) image
  const size_t group = (size_t)nphase * taps * unit;
  img.assign((size_t)li.nsplit * group, 0);
  for (int tap = 0; tap < taps; ++tap)
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci) {
        const float w = oihw[((size_t)co * cin + ci) * taps + tap] * wscale;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const int sp = co / nout, col = co % nout;
        const int ph = ci / li.cstage, cil = ci % li.cstage;
        const size_t base = (size_t)sp * group + ((size_t)ph * taps + tap) * unit;
        auto at = [&](int row) { return base + ((size_t)(cil >> 3) * (2 * li.np >> 3) + (row >> 3)) * 128 + (size_t)(row & 7) * 16 + (size_t)(cil & 7) * 2; };
        std::memcpy(&img[at(col)], &hi, 2);
        std::memcpy(&img[at(li.np + col)], &lo, 2);
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
