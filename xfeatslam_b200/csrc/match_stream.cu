// match_stream.cu -- brute-force 64-D descriptor matching (best / second-best per row) as a streaming tcgen05 GEMM.
//
// Output contract: the reference's INTEGER distance (ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:2242-2250:
// int(float(||a-b||^2) * 512), fp32 subtract + fp64 accumulate in index order) and its best / second-best scan
// (src/ORBmatcher.cc:476-486: strict '<', ascending index), bit-exact w.r.t. oracle/matcher_oracle.c.  ||a-b||^2 has no GEMM
// form that rounds like that, so the tensor cores FILTER and the few survivors are VERIFIED exactly.
//
//   * ms_prep_kernel: every descriptor becomes ONE fp16 operand row with the norm folded into the contraction:
//       as a ROW   : [ a_0 .. a_63 | 1 1 1 0 ... ]            (the constant tail is written straight into tensor memory)
//       as a COLUMN: [ b_0 .. b_63 | p1 p2 p3 0 ... ],  p1 + p2 + p3 = -|b|^2 / 2  (three fp16 pieces, residual 2^-33)
//     so the K = 80 GEMM yields  u_ij = a_i.b_j - |b_j|^2 / 2  directly and  t_ij = 512 |a_i|^2 - 1024 u_ij  ~  512 d_ij
//     with  |t - 512 float(d)| <= e_i = 1.05 sqrt(|a_i|^2 max|b|^2) + small  (fp16 rounding of both operands, rigorous;
//     a row with |x|^2 >= 1e5 does not fit the fp16 images -- its candidates are then ALL verified exactly, see `wild`).  Images are in the canonical K-major
//     no-swizzle UMMA layout (16-byte K chunks 2048 B apart, 8-row groups 128 B apart): a 128-row block = one 20 KB bulk copy.
//   * ms_kernel, one CTA per (128 rows, frame pair): the row block is copied ONCE into tensor memory (tcgen05.st, columns
//     384..423) and used as the A operand from there (tcgen05.mma with A in TMEM), so each 128 x 128 x 16 MMA reads only its
//     4 KB column slice from shared memory.  The column blocks stream through a 7-stage ring of bulk copies into three
//     128-column fp32 accumulators; 5 UTCHMMA (kind::f16) per accumulator.  The columns are streamed TWICE:
//       pass 1  epilogue = one 3-input fmax per element: the maximum of every 32-column slice (kept in shared memory, fp16
//               rounded up) and the row's two largest slice maxima.  The second-largest slice maximum is attained by a column
//               other than the one of the largest, so T2 = 512|a|^2 - 1024 * (it) bounds the row's second-smallest t.
//               Two columns have exact D <= T2 + e, hence every member of the exact top-2 has t < T2 + 2e + 1: the
//               threshold is FINAL after pass 1 (with a finite init_dist -- 256 in SearchByBoW -- it is also capped at init + e).
//       pass 2  only slices whose recorded maximum passes the threshold are read back from tensor memory (43 % at
//               4096 x 4096); a branch-free compare builds a 32-bit survivor mask per row and slice, survivors (3.6 per row on
//               XFeat descriptors) go to a per-warp queue.
//     Verification: a warp drains its queue 32 survivors at a time, one per lane, so the 64-step fp64 chain of the exact
//     distance is paid once per 32 survivors; results enter the row's exact top-2 with two 64-bit shared-memory atomicMin on
//     keys (D << 32 | column), i.e. the total order (distance, index) = the reference's scan order.
//     With vocabulary-node gating (group ids) pass 1 is skipped: the threshold is init + e alone.
//   * Column-wise best (mutual-NN checks) = the same kernel on the transposed problem (distances are symmetric bit for bit).
//
// Warp roles (448 threads): warp 0 = loader, warp 1 = TMEM allocator + MMA issuer (ONE lane elected with elect.sync runs the
// whole issue loop: barrier polls, UTCHMMA, UTCBAR), warps 2..13 = epilogue, three groups of four warps (one per TMEM lane
// quadrant); group g owns accumulator g, i.e. the tiles g, g + 3, ...  -- every accumulator barrier is followed phase by phase
// by ONE group (mbarrier parity waits alias when a waiter skips phases).
// Measured (tools/tmem_bench.cu, B200): tcgen05.ld 32x32b.x32 + wait = 48 clk for one warp, >= 900 B/clk/SM with 16 warps;
// kind::f16 M128 N128 K16 = 64 clk per MMA from shared or tensor memory: neither is what bounds this kernel -- the
// per-tile barrier handshakes of the issuing lane are (profiles/).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "tc_ptx.cuh"
#include "xfb_internal.h"

namespace xfb {

constexpr int MS_ROWS = 128;                       // rows (or columns) per operand block
constexpr int MS_ACC = 3;                          // accumulator stages (3 x 128 TMEM columns; the row operand lives in columns 384..423)
constexpr uint32_t MS_A_COL = 384;
constexpr uint32_t MS_LBO = 2048, MS_SBO = 128;    // bytes between 16-byte K chunks / between 8-row groups
constexpr int MS_BLK_BYTES = 10 * 2048;            // + the two K chunks of the column tail
constexpr int MS_STAGES = 7;                       // column blocks in flight
constexpr int MS_REC_SLICES = 128;                 // pass-1 slice maxima kept for up to 128 slices (4096 columns)
constexpr int MS_PARTS = 4;                        // 32-column slices per block = epilogue threads per row
constexpr int MS_EPI_WARPS = 4 * MS_ACC;           // one group of 4 warps (the 4 TMEM lane quadrants) per accumulator stage
constexpr int MS_THREADS = 64 + 32 * MS_EPI_WARPS;
constexpr int MS_QCAP = 128;                       // survivor queue entries per epilogue warp
constexpr float MS_PAD_AUG = -60000.0f;            // column tail of padded columns: u = -60000, never a candidate
// kind::f16: D = F32 (bit 4), A = B = F16 (format 0), both K-major, N = 128, M = 128
constexpr uint32_t MS_IDESC = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(MS_IDESC), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (lane = row, 16 fp16 of K = 8 consecutive 32-bit columns), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(MS_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"
      "%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]),
      "r"(u[13]), "r"(u[14]), "r"(u[15]), "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]), "r"(u[24]),
      "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* u) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]),
               "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t ms_desc(uint32_t smem_addr) { return umma_desc_kmajor(smem_addr, MS_LBO, MS_SBO); }

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float* v) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the same wait, tied to the 32 destination registers of an earlier tcgen05.ld so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait_dep(float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]), "+r"(u[8]), "+r"(u[9]), "+r"(u[10]),
                 "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15]), "+r"(u[16]), "+r"(u[17]), "+r"(u[18]), "+r"(u[19]), "+r"(u[20]),
                 "+r"(u[21]), "+r"(u[22]), "+r"(u[23]), "+r"(u[24]), "+r"(u[25]), "+r"(u[26]), "+r"(u[27]), "+r"(u[28]), "+r"(u[29]), "+r"(u[30]),
                 "+r"(u[31])
               :
               : "memory");
}

// ---- operand images ---------------------------------------------------------------------------------------
// One thread per (row, 4 consecutive k).  Rows >= n (per set) are zero-filled, with the padded-column tail.
__global__ void __launch_bounds__(256) ms_prep_kernel(const float* desc, size_t set_stride, const int32_t* n_dev, int n_host, int rows_padded,
                                                      unsigned char* img, size_t img_set_bytes, float* nrm, float* nrm_max) {
  const int set = blockIdx.y;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;   // row * 16 + kq
  const int row = g >> 4, kq = g & 15;
  if (row >= rows_padded) return;
  const int n = n_dev ? min(n_host, n_dev[set]) : n_host;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < n) v = *reinterpret_cast<const float4*>(desc + (size_t)set * set_stride + (size_t)row * 64 + kq * 4);
  const int blk = row >> 7, r = row & 127;
  unsigned char* base = img + (size_t)set * img_set_bytes + (size_t)blk * MS_BLK_BYTES;
  // element (r, k): chunk k / 8 at chunk * 2048, row at (r / 8) * 128 + (r % 8) * 16, 2 bytes per element
  const size_t off = (size_t)(kq >> 1) * MS_LBO + (size_t)(r >> 3) * MS_SBO + (size_t)(r & 7) * 16 + (size_t)(kq & 1) * 8;
  const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
  uint2 pk;
  pk.x = *reinterpret_cast<const unsigned int*>(&h01);
  pk.y = *reinterpret_cast<const unsigned int*>(&h23);
  *reinterpret_cast<uint2*>(base + off) = pk;
  // |a|^2: fp64 accumulate across the 16 threads of a row (lanes kq = 0..15 are contiguous in a half warp)
  double s = (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
  if (kq == 0) {
    const float nf = (float)s;
    nrm[(size_t)set * rows_padded + row] = (row < n) ? nf : CUDART_INF_F;
    if (s > 0.0 && row < n) atomicMax(reinterpret_cast<unsigned int*>(nrm_max + set), __float_as_uint(nf));   // non-negative floats order as uints
    // column tail: -|b|^2/2 = p1 + p2 + p3 (fp16 pieces), then zeros; padded columns get a huge negative tail
    float x = (row < n) ? -0.5f * nf : MS_PAD_AUG;
    const __half p1 = __float2half_rn(x);
    x -= __half2float(p1);
    const __half p2 = __float2half_rn(x);
    x -= __half2float(p2);
    const __half p3 = __float2half_rn(x);
    uint4 tail;
    tail.x = (unsigned int)__half_as_ushort(p1) | ((unsigned int)__half_as_ushort(p2) << 16);
    tail.y = (unsigned int)__half_as_ushort(p3);
    tail.z = 0u; tail.w = 0u;
    const size_t roff = (size_t)(r >> 3) * MS_SBO + (size_t)(r & 7) * 16;
    *reinterpret_cast<uint4*>(base + 8 * MS_LBO + roff) = tail;
    *reinterpret_cast<uint4*>(base + 9 * MS_LBO + roff) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// exact ORBmatcher::DescriptorDistance of two fp32 rows (same op order as oracle/matcher_oracle.c)
__device__ __forceinline__ int ms_exact_distance(const float* arow, const float* brow) {
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

struct MsShared {   // bookkeeping part of the dynamic shared memory (after the operand tiles)
  uint64_t bar_a, bar_full[MS_STAGES], bar_empty[MS_STAGES], bar_accf[MS_ACC], bar_acce[MS_ACC];
  uint32_t tmem;
};

__device__ __forceinline__ void ms_top2_u32(unsigned int* k1, unsigned int* k2, unsigned int key) {
  const unsigned int old = atomicMin(k1, key);
  atomicMin(k2, max(old, key));
}
__device__ __forceinline__ void ms_top2_u64(unsigned long long* k1, unsigned long long* k2, unsigned long long key) {
  const unsigned long long old = atomicMin(k1, key);
  atomicMin(k2, max(old, key));
}

// exact distance of one (row, column) -> the row's exact top-2 in the total order (D, column)
__device__ __noinline__ void ms_verify(const float* arow, const float* brow, unsigned long long* k1, unsigned long long* k2, unsigned int init_u, int j) {
  const int D = ms_exact_distance(arow, brow);
  if ((unsigned int)D < init_u) ms_top2_u64(k1, k2, ((unsigned long long)(unsigned int)D << 32) | (unsigned int)j);
}

// max of 32 accumulator values (3-input max tree)
__device__ __forceinline__ float ms_max32(const float* v) {
  float mg[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) mg[g] = fmaxf(fmaxf(v[4 * g], v[4 * g + 1]), fmaxf(v[4 * g + 2], v[4 * g + 3]));
  return fmaxf(fmaxf(fmaxf(mg[0], mg[1]), fmaxf(mg[2], mg[3])), fmaxf(fmaxf(mg[4], mg[5]), fmaxf(mg[6], mg[7])));
}

// debug: cycles spent inside a wait, accumulated into a.ms_counters[slot] by CTA (0, 0) only
#define MS_TIMED_WAIT(slot, stmt)                                                    \
  do {                                                                               \
    if (dbg) { const long long _t = clock64(); stmt; if ((threadIdx.x & 31) == 0) atomicAdd(a.ms_counters + (slot), (unsigned long long)(clock64() - _t)); } \
    else { stmt; }                                                                   \
  } while (0)

template <bool GROUPED>
__global__ void __launch_bounds__(MS_THREADS, 1) ms_kernel(const MatchTcArgs a) {
  const bool dbg = a.ms_counters != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
  const long long t_start = clock64();
  constexpr int ACC = MS_ACC;
  constexpr int GROUPS = MS_EPI_WARPS / 4;         // epilogue warp groups; tile k uses accumulator k % ACC and belongs to group k % GROUPS
  static_assert(GROUPS == ACC, "each accumulator barrier must be followed phase by phase by ONE group (parity waits alias otherwise)");
  constexpr int NROW = MS_ROWS;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sB0 = smem_raw;                                        // MS_STAGES x 20 KB
  __half* sRec = reinterpret_cast<__half*>(sB0 + MS_STAGES * MS_BLK_BYTES);  // [MS_REC_SLICES][128] slice maxima of pass 1, rounded UP to fp16
  MsShared* sh = reinterpret_cast<MsShared*>(sRec + MS_REC_SLICES * NROW);
  unsigned long long* sK1 = reinterpret_cast<unsigned long long*>(sh + 1);   // [NROW] exact best key
  unsigned long long* sK2 = sK1 + NROW;                                      // [NROW] exact second key
  unsigned int* sT1 = reinterpret_cast<unsigned int*>(sK2 + NROW);           // [NROW] approximate smallest t (float bits, >= 0)
  unsigned int* sT2 = sT1 + NROW;                                            // [NROW] approximate second-smallest t
  uint32_t* sQ = reinterpret_cast<uint32_t*>(sT2 + NROW);                    // [MS_EPI_WARPS][MS_QCAP] survivors: column | row << 24

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y;
  const int setA = a.pairs[2 * pair], setB = a.pairs[2 * pair + 1];
  const int nA = a.nA_dev ? min(a.nA_host, a.nA_dev[setA]) : a.nA_host;
  const int nB = a.nB_dev ? min(a.nB_host, a.nB_dev[setB]) : a.nB_host;
  const int row0 = blockIdx.x * NROW;
  const int nblk = (row0 < nA) ? (nB + MS_ROWS - 1) / MS_ROWS : 0;      // column blocks
  // Pass 1 (threshold) + pass 2 (survivors).  With vocabulary-node gating the threshold is init + e alone: one pass.
  const int passes = GROUPED ? 1 : 2;
  const int total = passes * nblk;                                      // accumulator tiles this CTA streams
  const unsigned char* imgA = reinterpret_cast<const unsigned char*>(a.imgA) + (size_t)setA * a.img_stride_A + (size_t)blockIdx.x * MS_BLK_BYTES;
  const unsigned char* imgB = reinterpret_cast<const unsigned char*>(a.imgB) + (size_t)setB * a.img_stride_B;
  const bool need2 = a.second_dist != nullptr;

  // ---- setup ----
  if (threadIdx.x < NROW) {
    const int r = threadIdx.x;
    const unsigned long long k0 = ((unsigned long long)(unsigned int)a.init << 32) | 0xffffffffull;
    sK1[r] = k0; sK2[r] = k0;
    sT1[r] = 0x7f800000u; sT2[r] = 0x7f800000u;
  }
  if (threadIdx.x == 0) {
    mbar_init(&sh->bar_a, 4);
    for (int s = 0; s < MS_STAGES; ++s) { mbar_init(&sh->bar_full[s], 1); mbar_init(&sh->bar_empty[s], 1); }
    for (int s = 0; s < ACC; ++s) { mbar_init(&sh->bar_accf[s], 1); mbar_init(&sh->bar_acce[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&sh->tmem, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem;

  if (warp == 0) {
    // ===== loader (one elected lane) =====
    if (elect_one_sync() && total > 0) {
      int c = 0;
      for (int k = 0; k < total; ++k) {
        const int s = k % MS_STAGES;
        if (k >= MS_STAGES) MS_TIMED_WAIT(7, mbar_wait(&sh->bar_empty[s], ((k / MS_STAGES) - 1) & 1));
        if (a.ms_mode & 16) mbar_arrive(&sh->bar_full[s]);   // timing experiment: no column-block copies
        else {
          mbar_expect_tx(&sh->bar_full[s], MS_BLK_BYTES);
          bulk_g2s(sB0 + (size_t)s * MS_BLK_BYTES, imgB + (size_t)c * MS_BLK_BYTES, MS_BLK_BYTES, &sh->bar_full[s]);
        }
        if (++c == nblk) c = 0;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: ONE elected lane runs the whole loop (barrier polls, tcgen05.mma, tcgen05.commit).  Under an
    // elect.sync predicate the tcgen05 instructions compile to single UTCHMMA / UTCBAR (under `lane == 0` the compiler
    // wraps each in a serialisation loop).  The row operand sits in tensor memory (written once by epilogue group 0), so
    // a 128 x 128 x 16 MMA reads only its 4 KB column slice from shared memory. =====
    const bool leader = elect_one_sync();
    if (leader && total > 0) {
      mbar_wait(&sh->bar_a, 0);
      tc_fence_after();
      const long long t_mma0 = clock64();
      const uint64_t db0 = ms_desc(smem_u32(sB0));
      constexpr uint64_t KSTEP = (2u * MS_LBO) >> 4;          // 16 fp16 = 2 K chunks, in descriptor address units (16 B)
      constexpr uint64_t BSTAGE = (uint64_t)MS_BLK_BYTES >> 4;
      const uint32_t ta = tmem_base + MS_A_COL;               // 8 columns per 16 fp16 of K
      const bool no_mma = (a.ms_mode & 4) != 0;               // (timing experiment)
      int s = 0, acc = 0;
      uint32_t sph = 0, aph = 0;
      long long c_full = 0, c_acce = 0, c_issue = 0, c_commit = 0;   // debug: cycles per section (registers, written once)
#pragma unroll 1
      for (int k = 0; k < total; ++k) {
        long long t0 = dbg ? clock64() : 0, t1;
        if (k >= ACC) mbar_wait2(&sh->bar_full[s], sph, &sh->bar_acce[acc], aph ^ 1u);   // column block landed AND accumulator drained
        else mbar_wait(&sh->bar_full[s], sph);
        if (dbg) { t1 = clock64(); c_full += t1 - t0; t0 = t1; }
        tc_fence_after();
        if (dbg) { t1 = clock64(); c_acce += t1 - t0; t0 = t1; }
        const uint64_t db = db0 + (uint64_t)s * BSTAGE;
        const uint32_t d = tmem_base + (uint32_t)acc * 128u;
        if (!no_mma) {
          umma_f16_ts(d, ta, db, 0u);
          umma_f16_ts(d, ta + 8u, db + KSTEP, 1u);
          umma_f16_ts(d, ta + 16u, db + 2 * KSTEP, 1u);
          umma_f16_ts(d, ta + 24u, db + 3 * KSTEP, 1u);
          umma_f16_ts(d, ta + 32u, db + 4 * KSTEP, 1u);       // row tail [1 1 1 0 ...] x column tail = -|b|^2 / 2
        }
        if (dbg) { t1 = clock64(); c_issue += t1 - t0; t0 = t1; }
        umma_commit(&sh->bar_empty[s]);   // smem stage may be refilled once these MMAs have read it
        umma_commit(&sh->bar_accf[acc]);  // accumulator complete
        if (dbg) { t1 = clock64(); c_commit += t1 - t0; }
        if (++s == MS_STAGES) { s = 0; sph ^= 1u; }
        if (++acc == ACC) { acc = 0; aph ^= 1u; }
      }
      if (dbg) {
        atomicAdd(a.ms_counters + 4, (unsigned long long)c_full); atomicAdd(a.ms_counters + 5, (unsigned long long)c_acce);
        atomicAdd(a.ms_counters + 14, (unsigned long long)c_issue); atomicAdd(a.ms_counters + 15, (unsigned long long)c_commit);
        atomicAdd(a.ms_counters + 6, (unsigned long long)(clock64() - t_mma0));
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: warp group g (4 warps = the 4 TMEM lane quadrants) owns accumulator stage g, i.e. the tiles
    // k = g, g + 4, ...; one handshake per 128 x 128 accumulator and warp =====
    const int ew = warp - 2;                         // (warps 2 .. 2 + MS_EPI_WARPS - 1)
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access (hardware: warp id % 4)
    const int group = ew >> 2;
    const int rs = quad * 32 + lane;                 // row inside the tile
    const float* rawA = a.rawA + (size_t)setA * a.raw_stride_A;
    const float* rawB = a.rawB + (size_t)setB * a.raw_stride_B;
    const unsigned int init_u = (unsigned int)a.init;
    const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16);

    if (group == 0 && total > 0) {
      // row operand -> tensor memory: lane = row, columns MS_A_COL .. +39 = 64 fp16 of the descriptor + the tail [1 1 1 0 ...].
      // In the image, K chunk ch of row r is the 16 bytes at ch * 2048 + r * 16: a warp reads 512 contiguous bytes per chunk.
      uint32_t w[32];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 x = *reinterpret_cast<const uint4*>(imgA + (size_t)ch * MS_LBO + (size_t)rs * 16);
        w[4 * ch] = x.x; w[4 * ch + 1] = x.y; w[4 * ch + 2] = x.z; w[4 * ch + 3] = x.w;
      }
      tmem_st32(tq + MS_A_COL, w);
      uint32_t t8[8] = {0x3C003C00u, 0x00003C00u, 0u, 0u, 0u, 0u, 0u, 0u};
      tmem_st8(tq + MS_A_COL + 32u, t8);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->bar_a);
    }

    uint32_t* q = sQ + (size_t)ew * MS_QCAP;         // this warp's survivor queue; its fill count `qn` is warp-uniform (a register)
    int qn = 0;
    // survivors are verified warp-wide, 32 at a time (one per lane): the 64-step fp64 chain of the exact distance is paid
    // once per batch instead of once per survivor
    auto drain = [&]() {
      __syncwarp();
      for (int i = lane; i < qn; i += 32) {
        const uint32_t ent = q[i];
        const int r = (int)(ent >> 24), j = (int)(ent & 0xffffffu);
        if (a.ms_counters && (a.ms_mode & 64)) atomicAdd(a.ms_counters + 1, 1ull);
        ms_verify(rawA + (size_t)(row0 + r) * 64, rawB + (size_t)j * 64, &sK1[r], &sK2[r], init_u, j);
      }
      qn = 0;
      __syncwarp();
    };

    const int row = row0 + rs;
    const bool ok = row < nA;
    float base = 0.f, margin = 0.f, cap = -CUDART_INF_F;
    bool wild = false;   // operands outside the fp16 filter's range (the column tail -|b|^2/2 must fit fp16): verify every column exactly
    if (ok) {
      const float na = a.nrmA[(size_t)setA * a.rows_padded_A + row], nbm = a.nrm_max_B[setB];
      wild = !(na < 1.0e5f && nbm < 1.0e5f);
      const float e = 1.05f * sqrtf(na * nbm) + 0.02f * (na + nbm + 1.0f);   // |t - 512 float(d)| <= e
      base = 512.0f * na;
      margin = 2.0f * e + 1.0f;
      cap = (a.init == 0x7fffffff) ? CUDART_INF_F : (float)a.init + e;
    }
    const int grp = (GROUPED && ok) ? a.gA[row] : 0;
    const bool use_rec = nblk * MS_PARTS <= MS_REC_SLICES;

    int k = group;
    if (passes == 2) {
      // ---- pass 1: slice maxima only.  The second-largest slice maximum of a row is attained by a column other than the
      // one of the largest, so T2 = base - 1024 * (second-largest slice maximum) bounds the row's second-smallest t. ----
      float m1 = -CUDART_INF_F, m2 = -CUDART_INF_F;
#pragma unroll 1
      for (; k < nblk; k += GROUPS) {
        const int acc = k % ACC;
        const uint32_t tm = tq + (uint32_t)acc * 128u;
        if (dbg && warp == 2 && lane == 0) MS_TIMED_WAIT(8, mbar_wait(&sh->bar_accf[acc], (uint32_t)(k / ACC) & 1u));
        else mbar_wait(&sh->bar_accf[acc], (uint32_t)(k / ACC) & 1u);
        __syncwarp();
        tc_fence_after();
        float va[32], vb[32];
        if (!(a.ms_mode & 8)) { tmem_ld32_nowait(tm, va); tmem_ld_wait_dep(va); }
#pragma unroll
        for (int part = 0; part < MS_PARTS; ++part) {
          float* cur = (part & 1) ? vb : va;
          float* nxt = (part & 1) ? va : vb;
          if (part + 1 < MS_PARTS && !(a.ms_mode & 8)) tmem_ld32_nowait(tm + (uint32_t)(part + 1) * 32u, nxt);   // in flight while `cur` is reduced
          const float m = ms_max32(cur);
          if (use_rec) sRec[(k * MS_PARTS + part) * NROW + rs] = __float2half_ru(m);   // rounded up: (stored > tau) never misses (m > tau)
          const float lo = fminf(m1, m);
          m1 = fmaxf(m1, m);
          m2 = fmaxf(m2, lo);
          if (part + 1 < MS_PARTS && !(a.ms_mode & 8)) tmem_ld_wait_dep(nxt);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->bar_acce[acc]);
      }
      if (ok) {
        if (m1 > -CUDART_INF_F) ms_top2_u32(&sT1[rs], &sT2[rs], __float_as_uint(fmaxf(fmaf(-1024.0f, m1, base), 0.f)));
        if (m2 > -CUDART_INF_F) ms_top2_u32(&sT1[rs], &sT2[rs], __float_as_uint(fmaxf(fmaf(-1024.0f, m2, base), 0.f)));
      }
      if (dbg && warp == 2 && lane == 0) atomicAdd(a.ms_counters + 9, (unsigned long long)(clock64() - t_start));   // end of this warp's pass 1
      asm volatile("bar.sync 1, %0;" ::"n"(32 * MS_EPI_WARPS) : "memory");   // every group's slice maxima are merged: thresholds are final
    }
    // t < thr  <=>  u > (base - thr) / 1024 ; the 1e-3 covers the re-association rounding
    const float thr = fminf(__uint_as_float(need2 ? sT2[rs] : sT1[rs]) + margin, cap);
    const float tau = ok ? ((thr == CUDART_INF_F) ? -CUDART_INF_F : (base - thr - 1e-3f) * (1.0f / 1024.0f)) : CUDART_INF_F;

    // ---- pass 2: collect the survivors (u > tau) ----
#pragma unroll 1
    for (; k < total; k += GROUPS) {
      const int c = k - (passes - 1) * nblk;
      const int acc = k % ACC;
      const uint32_t tm = tq + (uint32_t)acc * 128u;
      // which 32-column slices hold a survivor of one of this warp's rows (from the pass-1 slice maxima)
      uint32_t wneed = 0;
#pragma unroll
      for (int part = 0; part < MS_PARTS; ++part) {
        const bool need = (use_rec && passes == 2 && !wild) ? (__half2float(sRec[(c * MS_PARTS + part) * NROW + rs]) > tau) : ok;
        wneed |= __any_sync(0xffffffffu, need) ? (1u << part) : 0u;
      }
      if (a.ms_mode & 1) wneed = 0;                 // (timing experiment)
      mbar_wait(&sh->bar_accf[acc], (uint32_t)(k / ACC) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int part = 0; part < MS_PARTS; ++part) {
        if (!((wneed >> part) & 1u)) continue;
        float vv[32];
        tmem_ld32_nowait(tm + (uint32_t)part * 32u, vv);
        tmem_ld_wait_dep(vv);
        // branch-free bit mask of the columns with u > tau (rows without a survivor in this slice get 0)
        uint32_t mask = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) mask |= (vv[e] > tau) ? (1u << e) : 0u;
        if (wild) mask = 0xffffffffu;
        const int j0 = c * MS_ROWS + part * 32;
        if (j0 + 32 > nB) mask &= (nB > j0) ? (0xffffffffu >> (32 - (nB - j0))) : 0u;   // padded columns
        if (GROUPED) {   // vocabulary-node gating: drop the columns of other nodes
          uint32_t keep = 0;
          for (uint32_t mm = mask; mm; mm &= mm - 1u) { const int e = __ffs(mm) - 1; if (grp == a.gB[j0 + e]) keep |= 1u << e; }
          mask = keep;
        }
        // warp-uniform append (no atomics): every round, each lane with survivors left contributes its lowest one
        while (__any_sync(0xffffffffu, mask != 0u)) {
          if (qn > MS_QCAP - 32) drain();           // (uniform) room for one round
          const bool has = mask != 0u;
          const uint32_t bal = __ballot_sync(0xffffffffu, has);
          if (has) {
            const int e = __ffs(mask) - 1;
            mask &= mask - 1u;
            if (a.ms_counters && (a.ms_mode & 64)) atomicAdd(a.ms_counters + 0, 1ull);
            q[qn + __popc(bal & ((1u << lane) - 1u))] = (uint32_t)(j0 + e) | ((uint32_t)rs << 24);
          }
          qn += __popc(bal);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->bar_acce[acc]);
      if (qn >= MS_QCAP / 2) drain();
    }
    if (dbg && warp == 2 && lane == 0) atomicAdd(a.ms_counters + 10, (unsigned long long)(clock64() - t_start));   // end of pass 2
    drain();
    if (dbg && warp == 2 && lane == 0) atomicAdd(a.ms_counters + 11, (unsigned long long)(clock64() - t_start));   // end of the final drain
  }
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) { atomicAdd(a.ms_counters + 12, (unsigned long long)(clock64() - t_start)); atomicAdd(a.ms_counters + 13, 1ull); }
  if (threadIdx.x < NROW) {
    const int r = threadIdx.x, row = row0 + r;
    if (row < a.out_stride) {
      const bool row_ok = row < nA;
      const unsigned long long k1 = sK1[r], k2 = sK2[r];
      const size_t o = (size_t)pair * a.out_stride + row;
      if (a.best_idx) a.best_idx[o] = row_ok ? (int)(unsigned int)(k1 & 0xffffffffull) : -1;   // 0xffffffff = -1: none
      if (a.best_dist) a.best_dist[o] = row_ok ? (int)(k1 >> 32) : a.init;
      if (a.second_dist) a.second_dist[o] = row_ok ? (int)(k2 >> 32) : a.init;
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

constexpr size_t MS_SMEM = (size_t)MS_STAGES * MS_BLK_BYTES + (size_t)MS_REC_SLICES * MS_ROWS * 2 + sizeof(MsShared) +
                           (size_t)MS_ROWS * (8 + 8 + 4 + 4) + (size_t)MS_EPI_WARPS * MS_QCAP * 4 + 64;
static_assert(MS_SMEM <= 227 * 1024, "shared memory budget");

template <bool GROUPED>
static cudaError_t launch_ms(Ctx* c, const MatchTcArgs& a, int n_pairs) {
  auto kern = ms_kernel<GROUPED>;
  static unsigned long long attr_mask = 0;
  if (!((attr_mask >> c->device) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MS_SMEM);
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << c->device;
  }
  prof_begin(c, P_MATCH_TILE);
  kern<<<dim3(a.rows_padded_A / MS_ROWS, n_pairs), MS_THREADS, MS_SMEM, c->stream>>>(a);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

size_t ms_image_bytes(int rows_padded) { return (size_t)(rows_padded / MS_ROWS) * MS_BLK_BYTES; }

cudaError_t launch_ms_prep(Ctx* c, const float* desc, size_t set_stride, int n_sets, const int32_t* n_dev, int n_host, int rows_padded,
                           void* img, size_t img_set_bytes, float* nrm, float* nrm_max) {
  dim3 grid((rows_padded * 16 + 255) / 256, n_sets);
  cudaError_t e0 = cudaMemsetAsync(nrm_max, 0, (size_t)n_sets * 4, c->stream);
  if (e0 != cudaSuccess) return e0;
  prof_begin(c, P_MATCH_PREP);
  ms_prep_kernel<<<grid, 256, 0, c->stream>>>(desc, set_stride, n_dev, n_host, rows_padded, reinterpret_cast<unsigned char*>(img), img_set_bytes, nrm, nrm_max);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

// img_stride_A / img_stride_B of `a` are in BYTES here.
cudaError_t launch_match_stream(Ctx* c, const MatchTcArgs& a, int n_pairs, bool grouped) {
  return grouped ? launch_ms<true>(c, a, n_pairs) : launch_ms<false>(c, a, n_pairs);
}

}  // namespace xfb
