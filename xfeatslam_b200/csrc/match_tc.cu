// match_tc.cu -- the ALL-PAIRS distance table (xfb_distance_matrix) on the 5th-gen tensor cores (tcgen05 / TMEM).
//
// out[i][j] = ORBmatcher::DescriptorDistance(A_i, B_j) (src/ORBmatcher.cc:2242-2250: int(float(||a-b||^2) * 512), fp32 subtract +
// fp64 accumulate), bit-exact w.r.t. oracle/matcher_oracle.c, for every pair -- the table XFBmatcher::SearchForInitialization
// replays over.  (Best / second-best matching does not build this table: see match_stream.cu.)
//
//   0. match_prep_kernel: every descriptor is split into two bf16 pieces, a = a1 + a2 (+ 2^-18 |a|), and written as operand
//      images in the canonical K-major no-swizzle UMMA layout ([16-byte K chunk][8-row group][8 rows][16 B]; LBO = 2048 B,
//      SBO = 128 B), so a 128-row block is ONE contiguous bulk copy; |a|^2.
//   1. match_tc_kernel: dot(a_i, b_j) ~ a1.b1 + a1.b2 + a2.b1 (three bf16 UTCHMMA per 16 k; |error| < 3 * 2^-18 |a||b|), fp32
//      accumulators in TMEM;  t_ij = 512 (|a_i|^2 + |b_j|^2 - 2 dot) ~ 512 d_ij within +-eps(|a|,|b|) (match_eps).  Epilogue (one thread per
//      row = TMEM lane): D = floor(t) when t is farther than eps from an integer (92 % of the pairs), else the exact
//      fp64 re-evaluation from the original fp32 rows.
//
// Warp roles (576 threads): warp 0 = loader (cp.async.bulk + mbarrier complete_tx), warp 1 = TMEM allocator + tcgen05.mma
// issuer, warps 2..17 = epilogue (tcgen05.ld 32x32b; 4 warps per TMEM lane quadrant, each owning a 32-column slice).
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
// |t - 512*float(S)| <= eps(a, b): the bf16 two-piece split leaves |dot error| <= 3 * 2^-18 |a||b| (x 1024 = 0.0117 |a||b|), the fp32
// accumulation / re-association a few ulps of the operands' scale.  The bound is PER PAIR (it scales with the norms), so the
// floor() shortcut is sound for descriptors of any norm; when eps >= 0.5 every pair takes the exact path.
__device__ __forceinline__ float match_eps(float na, float nb) { return 0.03f * sqrtf(na * nb) + 2e-5f * 512.0f * (na + nb) + 0.005f; }
constexpr int TC_PARTS = 4;                       // 32-column slices per block = epilogue threads per row
constexpr int TC_EPI_WARPS = 4 * TC_PARTS;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_STAGES = 4;                      // B blocks in flight (main pass)
constexpr int TC_ACC = 4;                         // TMEM accumulator stages of the main pass (4 x 128 columns = all of TMEM)
constexpr int TC_NB_SMEM = 4096;                  // column norms kept in shared memory when n2 <= this
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
  float* sNb = reinterpret_cast<float*>(bars + 20);                   // [TC_NB_SMEM] |b|^2 of every column (when it fits)

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
    const float na = row_ok ? a.nrmA[(size_t)setA * a.rows_padded_A + row] : 0.f;
    const float base = 512.0f * na;
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
      // u = dot - |b|^2 / 2
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 nb = *reinterpret_cast<const float4*>(nrmB + j0 + 4 * g);
        v[4 * g + 0] = fmaf(-0.5f, nb.x, v[4 * g + 0]); v[4 * g + 1] = fmaf(-0.5f, nb.y, v[4 * g + 1]);
        v[4 * g + 2] = fmaf(-0.5f, nb.z, v[4 * g + 2]); v[4 * g + 3] = fmaf(-0.5f, nb.w, v[4 * g + 3]);
      }
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
            const float eps = match_eps(na, nrmB[j]);           // NaN / inf (huge norms) fail both comparisons -> exact path
            int D;
            if (fr > eps && fr < 1.0f - eps && t > 0.5f) D = (int)f;
            else D = exact_distance(arow, rawB + (size_t)j * 64);
            a.matrix[(size_t)row * nB + j] = D;
            if (a.dbg_maxerr) dbg_max = fmaxf(dbg_max, fabsf(t - exact_scaled(arow, rawB + (size_t)j * 64)));
        }
      }
    }
    if (a.dbg_maxerr) {
      atomicMax(reinterpret_cast<unsigned int*>(a.dbg_maxerr), __float_as_uint(dbg_max));   // non-negative floats order as uints
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_ACC * 128u);
  }
}

constexpr size_t TC_SMEM = (size_t)(1 + TC_STAGES) * TC_BLK_BYTES + 20 * 8 + TC_NB_SMEM * 4 + 64;

cudaError_t launch_matrix_tc(Ctx* c, const MatchTcArgs& a, int row_tiles) {
  static unsigned long long attr_mask = 0;
  if (!((attr_mask >> c->device) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM);
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << c->device;
  }
  prof_begin(c, P_DIST_MATRIX);
  match_tc_kernel<<<dim3(row_tiles, 1), TC_THREADS, TC_SMEM, c->stream>>>(a);
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


}  // namespace xfb
