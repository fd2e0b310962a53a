// match_tc.cu -- brute-force 64-D descriptor matching on the 5th-gen tensor cores (tcgen05 / TMEM).
//
// What must come out is the reference's INTEGER distance (ORBmatcher::DescriptorDistance,
// src/ORBmatcher.cc:2242-2250: int(float(||a-b||^2) * 512), fp32 subtract + fp64 accumulate) and the
// best / second-best scan of src/ORBmatcher.cc:476-486 (strict '<', ascending index), bit-exact w.r.t.
// oracle/matcher_oracle.c.  ||a-b||^2 has no GEMM form that rounds like that, so the matcher is a
// BOUND pass + FILTER pass + EXACT FIX-UP:
//
//   0. match_prep_kernel: every descriptor is split into two bf16 pieces, a = a1 + a2 (+ 2^-18 |a|), and
//      written as operand images in the canonical K-major no-swizzle UMMA layout ([16-byte K chunk][8-row
//      group][8 rows][16 B]; LBO = 2048 B, SBO = 128 B), so a 128-row block is ONE contiguous bulk copy; |a|^2.
//   1. match_bound_kernel (bf16 GEMM a1.b1, every 2nd column block, branch-free float top-2 per row): an upper
//      bound of each row's second-best distance.  Without it a running threshold sees ~2 ln(n) "records" per row.
//   2. match_tc_kernel: dot(a_i, b_j) ~ a1.b1 + a1.b2 + a2.b1 (three bf16 UTCHMMA per 16 k; |error| < 3 * 2^-18
//      |a||b|), fp32 accumulators in TMEM;  t_ij = 512 (|a_i|^2 + |b_j|^2 - 2 dot) ~ 512 d_ij within +-MATCH_EPS.
//      Epilogue (one thread per row = TMEM lane, columns ascending = the reference's scan): a pair can only
//      change (best, second) if t_ij < bound; for those few, D = floor(t) when t is farther than MATCH_EPS from
//      an integer, else the exact fp64 re-evaluation from the original fp32 rows.
//
// Warp roles (576 threads): warp 0 = loader (cp.async.bulk + mbarrier complete_tx), warp 1 = TMEM allocator +
// single-thread tcgen05.mma issuer, warps 2..17 = epilogue (tcgen05.ld 32x32b; 4 warps per TMEM lane quadrant,
// each owning a 32-column slice of every 128-column block, merged at the end in the total order (distance,
// index)).  The hot loop is branch-free: u = dot - |b|^2/2, a max over the 32 columns, ONE compare against tau.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "tc_ptx.cuh"
#include "xfb_internal.h"

namespace xfb {

constexpr int TC_ROWS = 128;                      // rows (or columns) per operand block
constexpr int TC_PIECE_BYTES = TC_ROWS * 64 * 2;  // one bf16 piece image of a block = 16 KB
constexpr int TC_BLK_BYTES = 2 * TC_PIECE_BYTES;  // [a1 | a2] = 32 KB
constexpr int TC_BLK_STRIDE = TC_BLK_BYTES / 4;   // in floats (the image buffers are addressed as float*)
constexpr float MATCH_EPS = 0.04f;                // bound on |t - 512*float(S)| (rigorous ~0.013 + fp32 accumulation; measured, see xfb_debug_match_error)
constexpr float MATCH_BF16_ERR = 4.2f;            // bound pass: |t_bf16 - 512 d| <= 1024 * 2^-8 * |a||b| (+ slack)
constexpr int TC_PARTS = 4;                       // 32-column slices per block = epilogue threads per row
constexpr int TC_EPI_WARPS = 4 * TC_PARTS;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_STAGES = 4;                      // B blocks in flight (main pass)
constexpr int TC_ACC = 4;                         // TMEM accumulator stages of the main pass (4 x 128 columns = all of TMEM)
constexpr int TC_NB_SMEM = 4096;                  // column norms kept in shared memory when n2 <= this
constexpr int TCB_STAGES = 4;                     // (bound pass)
constexpr int TCB_SKIP = 2;                       // bound pass visits every 2nd column block
constexpr uint32_t TC_LBO = 2048, TC_SBO = 128;
constexpr uint32_t TC_IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // kind::f16: BF16 x BF16 -> F32, M = N = 128
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) { return umma_desc_kmajor(smem_addr, TC_LBO, TC_SBO); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC_BF16), "r"(accumulate)
      : "memory");
}

// ---- operand images ---------------------------------------------------------------------------------------
// One thread per (row, 4 consecutive k).  Rows >= n (per set) are zero-filled up to the padded row count.
__global__ void __launch_bounds__(256) match_prep_kernel(const float* desc, size_t set_stride, const int32_t* n_dev, int n_host,
                                                         int rows_padded, float* img, size_t img_set_stride, float* nrm, float* nrm_max) {
  const int set = blockIdx.y;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;   // row * 16 + kq
  const int row = g >> 4, kq = g & 15;
  if (row >= rows_padded) return;
  const int n = n_dev ? min(n_host, n_dev[set]) : n_host;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < n) v = *reinterpret_cast<const float4*>(desc + (size_t)set * set_stride + (size_t)row * 64 + kq * 4);
  // a = a1 + a2 + O(2^-18 |a|): a1 = bf16(a), a2 = bf16(a - a1)   (a - a1 is exact in fp32)
  const __nv_bfloat16 p1x = __float2bfloat16_rn(v.x), p1y = __float2bfloat16_rn(v.y), p1z = __float2bfloat16_rn(v.z), p1w = __float2bfloat16_rn(v.w);
  const __nv_bfloat16 p2x = __float2bfloat16_rn(v.x - __bfloat162float(p1x)), p2y = __float2bfloat16_rn(v.y - __bfloat162float(p1y));
  const __nv_bfloat16 p2z = __float2bfloat16_rn(v.z - __bfloat162float(p1z)), p2w = __float2bfloat16_rn(v.w - __bfloat162float(p1w));
  const int blk = row >> 7, r = row & 127;
  unsigned int* base = reinterpret_cast<unsigned int*>(img + (size_t)set * img_set_stride + (size_t)blk * TC_BLK_STRIDE);
  // bf16 canonical layout: 8 elements per 16-byte chunk; 32-bit word index of element (r, k = 4 kq)
  const int widx = (((kq >> 1) * 16 + (r >> 3)) * 64 + (r & 7) * 8 + (kq & 1) * 4) >> 1;
  auto pack = [](__nv_bfloat16 lo, __nv_bfloat16 hi) { return (unsigned int)__bfloat16_as_ushort(lo) | ((unsigned int)__bfloat16_as_ushort(hi) << 16); };
  *reinterpret_cast<uint2*>(base + widx) = make_uint2(pack(p1x, p1y), pack(p1z, p1w));
  *reinterpret_cast<uint2*>(base + TC_PIECE_BYTES / 4 + widx) = make_uint2(pack(p2x, p2y), pack(p2z, p2w));
  // |a|^2: fp64 accumulate across the 16 threads of a row (lanes kq = 0..15 are contiguous in a half warp)
  double s = (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off, 16);
  if (kq == 0) {
    // padded rows get |b|^2 = +inf: as COLUMNS they then score u = dot - |b|^2/2 = -inf and can never be candidates
    nrm[(size_t)set * rows_padded + row] = (row < n) ? (float)s : CUDART_INF_F;
    if (s > 0.0) atomicMax(reinterpret_cast<unsigned int*>(nrm_max + set), __float_as_uint((float)s));   // non-negative floats order as uints
  }
}

// exact ORBmatcher::DescriptorDistance of two fp32 rows (same op order as oracle/matcher_oracle.c)
__device__ __noinline__ int exact_distance(const float* arow, const float* brow) {
  double s = 0.0;
#pragma unroll 4
  for (int kq = 0; kq < 16; ++kq) {
    const float4 x = *reinterpret_cast<const float4*>(arow + kq * 4);
    const float4 b = *reinterpret_cast<const float4*>(brow + kq * 4);
    float d;
    d = x.x - b.x; s = fma((double)d, (double)d, s);
    d = x.y - b.y; s = fma((double)d, (double)d, s);
    d = x.z - b.z; s = fma((double)d, (double)d, s);
    d = x.w - b.w; s = fma((double)d, (double)d, s);
  }
  return (int)(__double2float_rn(s) * 512.0f);
}
__device__ __noinline__ float exact_scaled(const float* arow, const float* brow) {
  double s = 0.0;
  for (int k = 0; k < 64; ++k) {
    const float d = arow[k] - brow[k];
    s = fma((double)d, (double)d, s);
  }
  return __double2float_rn(s) * 512.0f;
}

struct RowState { int b1, bidx, b2; float thr, tau, thr_fixed; };
__device__ __forceinline__ void row_state_refresh(RowState& st, float base) {
  st.thr = fminf((st.b2 == 0x7fffffff) ? CUDART_INF_F : (float)st.b2 + MATCH_EPS, st.thr_fixed);
  // t < thr  <=>  dot - |b|^2/2 > (base - thr)/1024 ; the extra 1e-3 covers the re-association rounding
  st.tau = (st.thr == CUDART_INF_F) ? -CUDART_INF_F : (base - st.thr - 1e-3f) * (1.0f / 1024.0f);
}

template <bool MATRIX, bool GROUPED>
__global__ void __launch_bounds__(TC_THREADS, 1) match_tc_kernel(const MatchTcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sA = smem_raw;                                        // [a1 16 KB | a2 16 KB]
  unsigned char* sB0 = smem_raw + TC_BLK_BYTES;                        // TC_STAGES x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB0 + TC_STAGES * TC_BLK_BYTES);
  uint64_t* bar_a = bars + 0;          // A tile landed
  uint64_t* bar_full = bars + 1;       // [4] B stage landed
  uint64_t* bar_empty = bars + 5;      // [4] B stage consumed by the tensor core
  uint64_t* bar_accf = bars + 9;       // [TC_ACC] accumulator ready
  uint64_t* bar_acce = bars + 13;      // [TC_ACC] accumulator drained by the epilogue
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 17);
  uint64_t* bar_nb = bars + 18;        // column norms landed
  int* sTop = reinterpret_cast<int*>(bars + 20);                      // [TC_PARTS][128][3] per-slice row results
  int* sBound = sTop + TC_PARTS * TC_ROWS * 3;                        // [128] best known upper bound of each row's second-best
  float* sNb = reinterpret_cast<float*>(sBound + TC_ROWS);            // [TC_NB_SMEM] |b|^2 of every column (when it fits)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y;
  const int setA = a.pairs[2 * pair], setB = a.pairs[2 * pair + 1];
  const int nA = a.nA_dev ? min(a.nA_host, a.nA_dev[setA]) : a.nA_host;
  const int nB = a.nB_dev ? min(a.nB_host, a.nB_dev[setB]) : a.nB_host;
  const int row0 = blockIdx.x * TC_ROWS;
  const int nblk = (row0 < nA) ? (nB + TC_ROWS - 1) / TC_ROWS : 0;    // column blocks to visit
  const float* imgA = a.imgA + (size_t)setA * a.img_stride_A + (size_t)blockIdx.x * TC_BLK_STRIDE;
  const float* imgB = a.imgB + (size_t)setB * a.img_stride_B;
  const bool nb_smem = nblk * TC_ROWS <= TC_NB_SMEM;

  if (threadIdx.x < TC_ROWS) sBound[threadIdx.x] = a.init;
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    mbar_init(bar_nb, 1);
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
    for (int s = 0; s < TC_ACC; ++s) { mbar_init(bar_accf + s, 1); mbar_init(bar_acce + s, TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(s_tmem, TC_ACC * 128u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===== loader =====
    if (lane == 0 && nblk > 0) {
      if (nb_smem) {   // the column norms are read by every epilogue thread for every block: keep them in shared memory
        mbar_expect_tx(bar_nb, (uint32_t)nblk * TC_ROWS * 4);
        bulk_g2s(sNb, a.nrmB + (size_t)setB * a.rows_padded_B, (uint32_t)nblk * TC_ROWS * 4, bar_nb);
      }
      mbar_expect_tx(bar_a, TC_BLK_BYTES);
      bulk_g2s(sA, imgA, TC_BLK_BYTES, bar_a);
      for (int c = 0; c < nblk; ++c) {
        const int s = c % TC_STAGES;
        if (c >= TC_STAGES) mbar_wait(bar_empty + s, ((c / TC_STAGES) - 1) & 1);
        mbar_expect_tx(bar_full + s, TC_BLK_BYTES);
        bulk_g2s(sB0 + (size_t)s * TC_BLK_BYTES, imgB + (size_t)c * TC_BLK_STRIDE, TC_BLK_BYTES, bar_full + s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && nblk > 0) {
      mbar_wait(bar_a, 0);
      const uint32_t a1 = smem_u32(sA), a2 = a1 + TC_PIECE_BYTES;
      for (int c = 0; c < nblk; ++c) {
        const int s = c % TC_STAGES, acc = c % TC_ACC;
        mbar_wait(bar_full + s, (c / TC_STAGES) & 1);
        if (c >= TC_ACC) mbar_wait(bar_acce + acc, ((c / TC_ACC) - 1) & 1);
        tc_fence_after();
        const uint32_t b1 = smem_u32(sB0 + (size_t)s * TC_BLK_BYTES), b2 = b1 + TC_PIECE_BYTES;
        const uint32_t d = tmem_base + (uint32_t)acc * 128u;
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint32_t ko = (uint32_t)k16 * 2u * TC_LBO;   // 16 bf16 = 2 k-chunks of 16 B
          umma_bf16(d, umma_desc(a1 + ko), umma_desc(b1 + ko), k16 > 0 ? 1u : 0u);
          umma_bf16(d, umma_desc(a1 + ko), umma_desc(b2 + ko), 1u);
          umma_bf16(d, umma_desc(a2 + ko), umma_desc(b1 + ko), 1u);
        }
        umma_commit(bar_empty + s);   // smem stage may be refilled once these MMAs have read it
        umma_commit(bar_accf + acc);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: TC_PARTS threads per row, each scanning its 32-column slice of every block =====
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access (hardware: warp id % 4)
    const int part = (warp - 2) >> 2;                // which 32-column slice of each 128-column block
    const int r = quad * 32 + lane;                  // row inside the tile
    const int row = row0 + r;
    const bool row_ok = row < nA;
    RowState st;
    st.b1 = a.init; st.b2 = a.init; st.bidx = -1;
    float base = 0.f;
    int grp = 0;
    st.thr_fixed = CUDART_INF_F;
    if (row_ok) {
      base = 512.0f * a.nrmA[(size_t)setA * a.rows_padded_A + row];
      if (GROUPED) grp = a.gA[row];
      // bound pass result: an upper bound of this row's second-best distance; +1 lets integer ties through
      if (!MATRIX && a.bound) st.thr_fixed = a.bound[(size_t)pair * a.rows_padded_A + row] + 1.0f + MATCH_EPS;
    }
    row_state_refresh(st, base);
    if (!row_ok) st.tau = CUDART_INF_F;              // padded rows never take the candidate path
    const float* nrmB = nb_smem ? sNb : a.nrmB + (size_t)setB * a.rows_padded_B;
    const float* rawB = a.rawB + (size_t)setB * a.raw_stride_B;
    const float* arow = a.rawA + (size_t)setA * a.raw_stride_A + (size_t)(row_ok ? row : 0) * 64;
    float dbg_max = 0.f;
    if (nb_smem && nblk > 0) mbar_wait(bar_nb, 0);
#pragma unroll 1
    for (int c = 0; c < nblk; ++c) {
      const int acc = c % TC_ACC;
      mbar_wait(bar_accf + acc, (c / TC_ACC) & 1);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * 128u + (uint32_t)part * 32u, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + acc);    // values are in registers: release the accumulator early
      const int j0 = c * TC_ROWS + part * 32;
      // hot path: u = dot - |b|^2/2, group maxima, one branch per 32 columns
      float mg[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 nb = *reinterpret_cast<const float4*>(nrmB + j0 + 4 * g);
        v[4 * g + 0] = fmaf(-0.5f, nb.x, v[4 * g + 0]); v[4 * g + 1] = fmaf(-0.5f, nb.y, v[4 * g + 1]);
        v[4 * g + 2] = fmaf(-0.5f, nb.z, v[4 * g + 2]); v[4 * g + 3] = fmaf(-0.5f, nb.w, v[4 * g + 3]);
        mg[g] = fmaxf(fmaxf(v[4 * g], v[4 * g + 1]), fmaxf(v[4 * g + 2], v[4 * g + 3]));
      }
      const float m = fmaxf(fmaxf(fmaxf(mg[0], mg[1]), fmaxf(mg[2], mg[3])), fmaxf(fmaxf(mg[4], mg[5]), fmaxf(mg[6], mg[7])));
      if (MATRIX) {
        if (row_ok) {
#pragma unroll 1
          for (int e = 0; e < 32; ++e) {
            const int j = j0 + e;
            if (j >= nB) break;
            float ue = v[0];
#pragma unroll
            for (int q = 1; q < 32; ++q) ue = (q == e) ? v[q] : ue;
            const float t = fmaf(-1024.0f, ue, base);
            const float f = floorf(t), fr = t - f;
            int D;
            if (fr > MATCH_EPS && fr < 1.0f - MATCH_EPS && t > 0.5f) D = (int)f;
            else D = exact_distance(arow, rawB + (size_t)j * 64);
            a.matrix[(size_t)row * nB + j] = D;
            if (a.dbg_maxerr) dbg_max = fmaxf(dbg_max, fabsf(t - exact_scaled(arow, rawB + (size_t)j * 64)));
          }
        }
      } else {
        // tighten the bound with what the other column slices of this row have found.  Ties must pass
        // (another slice may hold a higher column index with the same distance): shared bound is b2 + 1.
        const int shared_b2 = sBound[r];
        float tau = st.tau;
        if (shared_b2 != 0x7fffffff) tau = fmaxf(tau, (base - ((float)shared_b2 + 1.0f + MATCH_EPS) - 1e-3f) * (1.0f / 1024.0f));
        if (m > tau) {
          // candidate path (a few events per row per match): descend through the group maxima
          bool improved = false;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (mg[g] > tau) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int e = 4 * g + q;
                if (v[e] > tau) {
                  const int j = j0 + e;
                  if (j < nB && (!GROUPED || grp == a.gB[j])) {
                    const float t = fmaf(-1024.0f, v[e], base);     // = 512 (|a|^2 + |b|^2 - 2 dot)
                    const float f = floorf(t), fr = t - f;
                    int D;
                    if (fr > MATCH_EPS && fr < 1.0f - MATCH_EPS && t > 0.5f) D = (int)f;
                    else D = exact_distance(arow, rawB + (size_t)j * 64);
                    if (D < st.b1) { st.b2 = st.b1; st.b1 = D; st.bidx = j; improved = true; }
                    else if (D < st.b2) { st.b2 = D; improved = true; }
                  }
                }
              }
            }
          }
          if (improved) {
            row_state_refresh(st, base);
            if (st.b2 < shared_b2) atomicMin(&sBound[r], st.b2);
          }
        }
      }
    }
    if (!MATRIX) {
      sTop[(part * TC_ROWS + r) * 3 + 0] = st.b1;
      sTop[(part * TC_ROWS + r) * 3 + 1] = st.bidx;
      sTop[(part * TC_ROWS + r) * 3 + 2] = st.b2;
    } else if (a.dbg_maxerr) {
      atomicMax(reinterpret_cast<unsigned int*>(a.dbg_maxerr), __float_as_uint(dbg_max));   // non-negative floats order as uints
    }
  }
  tc_fence_before();
  __syncthreads();
  if (!MATRIX && threadIdx.x < TC_ROWS) {
    // merge the column slices of each row in the total order (distance, index) -- order independent, exact
    const int r = threadIdx.x, row = row0 + r;
    int b1 = sTop[r * 3], bidx = sTop[r * 3 + 1], b2 = sTop[r * 3 + 2];
#pragma unroll
    for (int p = 1; p < TC_PARTS; ++p) {
      const int o1 = sTop[(p * TC_ROWS + r) * 3], oi = sTop[(p * TC_ROWS + r) * 3 + 1], o2 = sTop[(p * TC_ROWS + r) * 3 + 2];
      if (b1 < o1 || (b1 == o1 && (unsigned)bidx <= (unsigned)oi)) { b2 = min(b2, o1); }
      else { b2 = min(o2, b1); b1 = o1; bidx = oi; }
    }
    if (row < a.out_stride) {
      const bool row_ok = row < nA;
      const size_t o = (size_t)pair * a.out_stride + row;
      if (a.best_idx) a.best_idx[o] = row_ok ? bidx : -1;
      if (a.best_dist) a.best_dist[o] = row_ok ? b1 : a.init;
      if (a.second_dist) a.second_dist[o] = row_ok ? b2 : a.init;
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_ACC * 128u);
  }
}

// ---- bound pass ----------------------------------------------------------------------------------------------
// A cheap first pass (bf16 GEMM a1.b1, branch-free float top-2 per row) that gives every row an upper bound on
// its SECOND-BEST distance.  bound = (approximate second-smallest t over the visited columns) + (rigorous bf16
// error): valid because at least two columns have a true distance below it.
__device__ __forceinline__ void top2_max_push(float& m1, float& m2, float u) {
  const float lo = fminf(m1, u);
  m1 = fmaxf(m1, u);
  m2 = fmaxf(m2, lo);
}

__global__ void __launch_bounds__(TC_THREADS, 2) match_bound_kernel(const MatchTcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sA = smem_raw;                                        // 16 KB bf16 tile (a1)
  unsigned char* sB0 = smem_raw + TC_PIECE_BYTES;                      // TCB_STAGES x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB0 + TCB_STAGES * TC_PIECE_BYTES);
  uint64_t* bar_a = bars + 0;
  uint64_t* bar_full = bars + 1;       // [4]
  uint64_t* bar_empty = bars + 5;      // [4]
  uint64_t* bar_accf = bars + 9;       // [2]
  uint64_t* bar_acce = bars + 11;      // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 13);
  uint64_t* bar_nb = bars + 14;
  float* sM = reinterpret_cast<float*>(bars + 16);                    // [TC_PARTS][128][2]
  float* sNb = sM + TC_PARTS * TC_ROWS * 2;                           // [TC_NB_SMEM]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y;
  const int setA = a.pairs[2 * pair], setB = a.pairs[2 * pair + 1];
  const int nA = a.nA_dev ? min(a.nA_host, a.nA_dev[setA]) : a.nA_host;
  const int nB = a.nB_dev ? min(a.nB_host, a.nB_dev[setB]) : a.nB_host;
  const int row0 = blockIdx.x * TC_ROWS;
  // Only every TCB_SKIP-th column block is visited: the second-smallest distance over a SUBSET of the columns
  // is still an upper bound of the second-smallest over all of them (slightly looser, half the work).
  const int nblk_all = (nB + TC_ROWS - 1) / TC_ROWS;
  const int nblk = (row0 < nA) ? (nblk_all + TCB_SKIP - 1) / TCB_SKIP : 0;
  const float* imgA = a.imgA + (size_t)setA * a.img_stride_A + (size_t)blockIdx.x * TC_BLK_STRIDE;
  const float* imgB = a.imgB + (size_t)setB * a.img_stride_B;
  const bool nb_smem = nblk_all * TC_ROWS <= TC_NB_SMEM;

  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    mbar_init(bar_nb, 1);
    for (int s = 0; s < TCB_STAGES; ++s) { mbar_init(bar_full + s, 1); mbar_init(bar_empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_accf + s, 1); mbar_init(bar_acce + s, TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(s_tmem, 256u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    if (lane == 0 && nblk > 0) {
      if (nb_smem) {
        mbar_expect_tx(bar_nb, (uint32_t)nblk_all * TC_ROWS * 4);
        bulk_g2s(sNb, a.nrmB + (size_t)setB * a.rows_padded_B, (uint32_t)nblk_all * TC_ROWS * 4, bar_nb);
      }
      mbar_expect_tx(bar_a, TC_PIECE_BYTES);
      bulk_g2s(sA, imgA, TC_PIECE_BYTES, bar_a);
      for (int c = 0; c < nblk; ++c) {
        const int s = c % TCB_STAGES;
        if (c >= TCB_STAGES) mbar_wait(bar_empty + s, ((c / TCB_STAGES) - 1) & 1);
        mbar_expect_tx(bar_full + s, TC_PIECE_BYTES);
        bulk_g2s(sB0 + (size_t)s * TC_PIECE_BYTES, imgB + (size_t)(c * TCB_SKIP) * TC_BLK_STRIDE, TC_PIECE_BYTES, bar_full + s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nblk > 0) {
      mbar_wait(bar_a, 0);
      const uint32_t a_addr = smem_u32(sA);
      for (int c = 0; c < nblk; ++c) {
        const int s = c % TCB_STAGES, acc = c & 1;
        mbar_wait(bar_full + s, (c / TCB_STAGES) & 1);
        if (c >= 2) mbar_wait(bar_acce + acc, ((c >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t b_addr = smem_u32(sB0 + (size_t)s * TC_PIECE_BYTES);
        const uint32_t d = tmem_base + (uint32_t)acc * 128u;
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          const uint32_t ko = (uint32_t)k16 * 2u * TC_LBO;   // 16 bf16 = 2 k-chunks of 16 B
          umma_bf16(d, umma_desc(a_addr + ko), umma_desc(b_addr + ko), k16 > 0 ? 1u : 0u);
        }
        umma_commit(bar_empty + s);
        umma_commit(bar_accf + acc);
      }
    }
  } else {
    const int quad = warp & 3, part = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const float* nrmB = nb_smem ? sNb : a.nrmB + (size_t)setB * a.rows_padded_B;
    if (nb_smem && nblk > 0) mbar_wait(bar_nb, 0);
    float m1a = -CUDART_INF_F, m2a = -CUDART_INF_F, m1b = -CUDART_INF_F, m2b = -CUDART_INF_F;
#pragma unroll 1
    for (int c = 0; c < nblk; ++c) {
      const int acc = c & 1;
      mbar_wait(bar_accf + acc, (c >> 1) & 1);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * 128u + (uint32_t)part * 32u, v);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + acc);
      const int j0 = (c * TCB_SKIP) * TC_ROWS + part * 32;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 nb = *reinterpret_cast<const float4*>(nrmB + j0 + 4 * g);   // +inf for padded columns -> u = -inf
        const float u0 = fmaf(-0.5f, nb.x, v[4 * g + 0]), u1 = fmaf(-0.5f, nb.y, v[4 * g + 1]);
        const float u2 = fmaf(-0.5f, nb.z, v[4 * g + 2]), u3 = fmaf(-0.5f, nb.w, v[4 * g + 3]);
        top2_max_push(m1a, m2a, u0); top2_max_push(m1b, m2b, u1);
        top2_max_push(m1a, m2a, u2); top2_max_push(m1b, m2b, u3);
      }
    }
    const float M1 = fmaxf(m1a, m1b), M2 = fmaxf(fminf(m1a, m1b), fmaxf(m2a, m2b));
    sM[(part * TC_ROWS + r) * 2] = M1;
    sM[(part * TC_ROWS + r) * 2 + 1] = M2;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < TC_ROWS) {
    const int r = threadIdx.x, row = row0 + r;
    float M1 = sM[r * 2], M2 = sM[r * 2 + 1];
#pragma unroll
    for (int p = 1; p < TC_PARTS; ++p) {
      const float o1 = sM[(p * TC_ROWS + r) * 2], o2 = sM[(p * TC_ROWS + r) * 2 + 1];
      M2 = fmaxf(fminf(M1, o1), fmaxf(M2, o2));
      M1 = fmaxf(M1, o1);
    }
    if (row < a.rows_padded_A) {
      float bnd = CUDART_INF_F;
      if (row < nA && M2 > -CUDART_INF_F) {
        const float na = a.nrmA[(size_t)setA * a.rows_padded_A + row];
        const float t2 = fmaf(-1024.0f, M2, 512.0f * na);                      // approximate second-smallest 512 d
        bnd = t2 + MATCH_BF16_ERR * sqrtf(na * a.nrm_max_B[setB]) + 0.5f;
      }
      a.bound[(size_t)pair * a.rows_padded_A + row] = bnd;
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256u);
  }
}

constexpr size_t TCB_SMEM = (size_t)(1 + TCB_STAGES) * TC_PIECE_BYTES + 16 * 8 + (size_t)TC_PARTS * TC_ROWS * 2 * 4 + TC_NB_SMEM * 4 + 64;
constexpr size_t TC_SMEM = (size_t)(1 + TC_STAGES) * TC_BLK_BYTES + 20 * 8 + (size_t)TC_PARTS * TC_ROWS * 3 * 4 + TC_ROWS * 4 + TC_NB_SMEM * 4 + 64;

template <bool MATRIX, bool GROUPED>
static cudaError_t launch_tc(Ctx* c, const MatchTcArgs& a, int row_tiles, int n_pairs, int tag) {
  auto kern = match_tc_kernel<MATRIX, GROUPED>;
  static unsigned long long attr_mask = 0;
  if (!((attr_mask >> c->device) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM);
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << c->device;
  }
  prof_begin(c, tag);
  kern<<<dim3(row_tiles, n_pairs), TC_THREADS, TC_SMEM, c->stream>>>(a);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

cudaError_t launch_match_prep(Ctx* c, const float* desc, size_t set_stride, int n_sets, const int32_t* n_dev, int n_host, int rows_padded,
                              float* img, size_t img_set_stride, float* nrm, float* nrm_max) {
  dim3 grid((rows_padded * 16 + 255) / 256, n_sets);
  cudaError_t e0 = cudaMemsetAsync(nrm_max, 0, (size_t)n_sets * 4, c->stream);
  if (e0 != cudaSuccess) return e0;
  prof_begin(c, P_MATCH_PREP);
  match_prep_kernel<<<grid, 256, 0, c->stream>>>(desc, set_stride, n_dev, n_host, rows_padded, img, img_set_stride, nrm, nrm_max);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

cudaError_t launch_match_bound(Ctx* c, const MatchTcArgs& a, int row_tiles, int n_pairs) {
  static unsigned long long attr_mask = 0;
  if (!((attr_mask >> c->device) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(match_bound_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TCB_SMEM);
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << c->device;
  }
  prof_begin(c, P_MATCH_BOUND);
  match_bound_kernel<<<dim3(row_tiles, n_pairs), TC_THREADS, TCB_SMEM, c->stream>>>(a);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

cudaError_t launch_match_tc(Ctx* c, const MatchTcArgs& a, int row_tiles, int n_pairs, bool grouped) {
  return grouped ? launch_tc<false, true>(c, a, row_tiles, n_pairs, P_MATCH_TILE) : launch_tc<false, false>(c, a, row_tiles, n_pairs, P_MATCH_TILE);
}
cudaError_t launch_matrix_tc(Ctx* c, const MatchTcArgs& a, int row_tiles) { return launch_tc<true, false>(c, a, row_tiles, 1, P_DIST_MATRIX); }

}  // namespace xfb
