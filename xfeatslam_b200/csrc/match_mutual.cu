// match_mutual.cu -- mutual brute-force matching (row-wise best / second-best AND column-wise best) of two 64-D descriptor
// sets in ONE pass over ONE streamed GEMM per pair, on the tcgen05 tensor cores.
//
// Output contract (unchanged, bit-exact w.r.t. oracle/matcher_oracle.c): the reference's INTEGER distance
// (ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:2242-2250: int(float(||a-b||^2) * 512)), its best / second-best scan per row
// (src/ORBmatcher.cc:476-486: strict '<', ascending index) and the column-wise argmin of the same matrix (the mutual-NN check of
// the commented-out ORBmatcher::match, src/ORBmatcher.cc:340-406).  As in match_stream.cu the tensor cores FILTER and the few
// survivors are VERIFIED with the exact arithmetic; what is new is that everything happens in one stream of the column set:
//
//   * Operand images: every descriptor x is ONE fp16 row of K = 80: [x_0 .. x_63 | tail], and carries BOTH tails:
//       as a ROW    vector: [ p1 p2 p3 1 1 1 0 0 ]        as a COLUMN vector: [ 1 1 1 p1 p2 p3 0 0 ],   p1 + p2 + p3 = -|x|^2 / 2
//     so the GEMM yields  u_ij = a_i.b_j - |a_i|^2/2 - |b_j|^2/2 = -||a_i - b_j||^2 / 2  directly:  512 d_ij ~ t_ij = -1024 u_ij
//     with |t - 512 float(d)| <= e = 1.05 |a||b| + small (fp16 rounding of both operands).  Larger u = smaller distance, for rows
//     and columns alike, no per-row / per-column offset.  Padded rows / columns carry a tail of -60000 and never win.
//   * Per (128 rows of A, pair) one CTA streams the 128-column blocks of B ONCE through a ring of bulk copies and issues, per
//     block, TWO accumulators: D1 = A_blk B_c^T (TMEM lane = row, columns = the block's columns) and D2 = B_c A_blk^T (lane =
//     column, columns = the CTA's rows) -- the same inner products, transposed by the tensor core, so that BOTH directions get the
//     cheap "one lane owns one line of 128 values" epilogue.  Two accumulator stages of 2 x 128 TMEM columns.
//   * Row direction (8 warps, two groups alternating blocks): running top-2 of the 32-column slice maxima per row (the
//     second-largest slice maximum is attained by another column than the largest, so it bounds the row's second-smallest t);
//     a column is a CANDIDATE when u > (running bound) - (2e + 1)/1024.  The bound only tightens; candidates are queued with the
//     slice maximum as an upper bound of their u and re-filtered against the FINAL bound before the exact verification.
//   * Column direction (8 warps), ONE pass per block: a row is queued when its value exceeds (the largest value known for the
//     column) - (2e + 1)/1024, "known" = what the pair's other CTAs have published so far in a global array (atomicMax; the CTAs walk
//     the column blocks in rotated order, so on average half of them have been there already) and this CTA's own rows up to the
//     current 32-row part.  Any such bound is <= the final column maximum: no candidate is lost.  At the end a CTA verifies only
//     the queued rows that are still within the margin of the FINAL maximum, and merges exact (distance, row) keys with a 64-bit
//     global atomicMin -- the total order (distance, index) of the reference's scan.  The last CTA of a pair writes the column
//     outputs and resets the scratch.
//   * Verification: 8 queued pairs at a time, four lanes per pair (all of a lane's row loads in flight at once).
//
// Warp roles (576 threads): warp 0 loader, warp 1 TMEM allocator + MMA issuer (one elected lane), warps 2-9 row direction,
// warps 10-17 column direction (TMEM lane quadrant = warp id % 4).  Vocabulary-node gated searches (group ids) and descriptor
// sets outside the fp16 range keep using match_stream.cu.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "tc_ptx.cuh"
#include "xfb_internal.h"

namespace xfb {

constexpr int MM_ROWS = 128;
constexpr int MM_CHUNKS = 11;                          // 8 data chunks (8 fp16 each) + row tail + column tail + zeros
constexpr int MM_BLK_BYTES = MM_CHUNKS * 2048;         // one 128-descriptor block: 22 KB, one bulk copy
constexpr uint32_t MM_LBO = 2048, MM_SBO = 128;        // bytes between 16-byte K chunks / between 8-row groups
constexpr int MM_STAGES = 4;                           // column blocks in flight
constexpr int MM_EPI_WARPS = 16;
constexpr int MM_THREADS = 64 + 32 * MM_EPI_WARPS;
constexpr int MM_Q1 = 512, MM_Q2 = 768;                // queue entries per row-direction / column-direction warp
constexpr float MM_PAD_TAIL = -60000.0f;
constexpr uint32_t MM_IDESC = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // kind::f16, F16 x F16 -> F32, M = N = 128

size_t mm_image_bytes(int rows_padded) { return (size_t)(rows_padded / MM_ROWS) * MM_BLK_BYTES; }

// ---- operand images: one CTA per 128-descriptor block.  The rows come in as coalesced float4 (16 lanes per row), the fp16 image is
// assembled in shared memory in its final K-major layout and leaves as ONE contiguous 22 KB run of 16-byte stores (the first version
// wrote 8 scattered bytes per thread: 51 us per 32 x 4096 descriptors, four times the byte time).  Rows >= n (per set) are zero with
// the padding tails.
__global__ void __launch_bounds__(256) mm_prep_kernel(const float* desc, size_t set_stride, const int32_t* n_dev, int n_host, int rows_padded,
                                                      unsigned char* img, size_t img_set_bytes, float* nrm, float* nrm_max) {
  __shared__ __align__(16) unsigned char simg[MM_BLK_BYTES];
  __shared__ unsigned int s_max;
  const int set = blockIdx.y, blk = blockIdx.x, t = threadIdx.x;
  const int n = n_dev ? min(n_host, n_dev[set]) : n_host;
  if (t == 0) s_max = 0u;
  __syncthreads();
  const int kq = t & 15;
  unsigned int my_max = 0u;
#pragma unroll 2
  for (int it = 0; it < MM_ROWS / 16; ++it) {
    const int r = it * 16 + (t >> 4), row = blk * MM_ROWS + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n) v = *reinterpret_cast<const float4*>(desc + (size_t)set * set_stride + (size_t)row * 64 + kq * 4);
    const size_t roff = (size_t)(r >> 3) * MM_SBO + (size_t)(r & 7) * 16;
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const unsigned int*>(&h01);
    pk.y = *reinterpret_cast<const unsigned int*>(&h23);
    *reinterpret_cast<uint2*>(simg + (size_t)(kq >> 1) * MM_LBO + roff + (size_t)(kq & 1) * 8) = pk;
    double s = (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, 16);
    if (kq == 0) {
      const float nf = (float)s;
      nrm[(size_t)set * rows_padded + row] = (row < n) ? nf : CUDART_INF_F;
      if (s > 0.0 && row < n) my_max = max(my_max, __float_as_uint(nf));
      float x = (row < n) ? -0.5f * nf : MM_PAD_TAIL;       // -|x|^2/2 = p1 + p2 + p3 (fp16 pieces, residual 2^-33)
      const __half p1 = __float2half_rn(x);
      x -= __half2float(p1);
      const __half p2 = __float2half_rn(x);
      x -= __half2float(p2);
      const __half p3 = __float2half_rn(x);
      const unsigned int p12 = (unsigned int)__half_as_ushort(p1) | ((unsigned int)__half_as_ushort(p2) << 16);
      const unsigned int p3u = (unsigned int)__half_as_ushort(p3);
      // row-role tail [p1 p2 p3 1 | 1 1 0 0], column-role tail [1 1 1 p1 | p2 p3 0 0]; 1.0 = 0x3C00
      *reinterpret_cast<uint4*>(simg + 8 * MM_LBO + roff) = make_uint4(p12, p3u | 0x3C000000u, 0x3C003C00u, 0u);
      *reinterpret_cast<uint4*>(simg + 9 * MM_LBO + roff) = make_uint4(0x3C003C00u, 0x00003C00u | (p12 << 16), (p12 >> 16) | (p3u << 16), 0u);
      *reinterpret_cast<uint4*>(simg + 10 * MM_LBO + roff) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (my_max) atomicMax(&s_max, my_max);
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(img + (size_t)set * img_set_bytes + (size_t)blk * MM_BLK_BYTES);
  const uint4* src = reinterpret_cast<const uint4*>(simg);
  for (int i = t; i < MM_BLK_BYTES / 16; i += 256) dst[i] = src[i];
  if (t == 0 && s_max) atomicMax(reinterpret_cast<unsigned int*>(nrm_max + set), s_max);
}

__device__ __forceinline__ void mm_umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(MM_IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mm_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
      "%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
        "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]),
        "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]),
        "=r"(u[31])
      : "r"(taddr));
  // the wait is tied to the 32 destination registers, so that no use of them can be scheduled above it
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]), "+r"(u[8]), "+r"(u[9]), "+r"(u[10]),
                 "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15]), "+r"(u[16]), "+r"(u[17]), "+r"(u[18]), "+r"(u[19]), "+r"(u[20]),
                 "+r"(u[21]), "+r"(u[22]), "+r"(u[23]), "+r"(u[24]), "+r"(u[25]), "+r"(u[26]), "+r"(u[27]), "+r"(u[28]), "+r"(u[29]), "+r"(u[30]),
                 "+r"(u[31])
               :
               : "memory");
}
__device__ __forceinline__ float mm_max32(const float* v) {
  float mg[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) mg[g] = fmaxf(fmaxf(v[4 * g], v[4 * g + 1]), fmaxf(v[4 * g + 2], v[4 * g + 3]));
  return fmaxf(fmaxf(fmaxf(mg[0], mg[1]), fmaxf(mg[2], mg[3])), fmaxf(fmaxf(mg[4], mg[5]), fmaxf(mg[6], mg[7])));
}
// order-preserving float -> unsigned (0 is below every value)
// Candidate mask of one 32-value slice: bit e = (v[e] > tau), possibly MORE bits than that (never fewer) when |tau| < 2^-60.
// The slice maximum costs 16 alu-pipe instructions; a compare + bit insert costs 2.5 more per value on the same pipe
// (FSETP, SEL, half an IADD3) while the fma pipe idles.  So MM_FMA_BITS of the 32 values take the fma pipe instead:
//   s = saturate(fma(v, 2^100, -tau * 2^100))     1.0 when v > tau, 0.0 otherwise   (FFMA.SAT)
//   acc = fma(s, 2^e, acc)                         acc starts at 2^23: integer sums < 2^23 are exact in the mantissa
// fma(v, 2^100, -tau 2^100) is (v - tau) * 2^100 with ONE rounding, so its sign is exact; it is >= 1 whenever v > tau because
// |tau| >= 2^-60 puts distinct values at least 2^-84 apart (tau closer to zero than that is replaced by -2^-60: a looser
// bound, a superset of candidates, all of which are verified exactly later); |v|, |tau| < 2^18 here, no overflow.
#ifndef XFB_MM_FMA_BITS
#define XFB_MM_FMA_BITS 21
#endif
constexpr int MM_FMA_BITS = XFB_MM_FMA_BITS;      // 0: every compare on the alu pipe
__device__ __forceinline__ uint32_t mm_mask32(const float* v, float tau) {
  constexpr float BIG = 0x1p100f;
  const float te = (fabsf(tau) < 0x1p-60f) ? -0x1p-60f : tau;
  const float nt = -te * BIG;                              // -inf / +inf pass through: everything / nothing is a candidate
  constexpr int H = (MM_FMA_BITS + 1) / 2;
  float acc0 = 8388608.0f, acc1 = 8388608.0f;
#pragma unroll
  for (int e = 0; e < H; ++e) acc0 = fmaf(__saturatef(fmaf(v[e], BIG, nt)), (float)(1u << e), acc0);
#pragma unroll
  for (int e = H; e < MM_FMA_BITS; ++e) acc1 = fmaf(__saturatef(fmaf(v[e], BIG, nt)), (float)(1u << (e - H)), acc1);
  uint32_t mk0 = 0u, mk1 = 0u;
#pragma unroll
  for (int e = MM_FMA_BITS; e < 32; e += 2) mk0 |= (v[e] > tau) ? (1u << e) : 0u;
#pragma unroll
  for (int e = MM_FMA_BITS + 1; e < 32; e += 2) mk1 |= (v[e] > tau) ? (1u << e) : 0u;
  return (__float_as_uint(acc0) & 0x007fffffu) | ((__float_as_uint(acc1) & 0x007fffffu) << H) | (mk0 | mk1);
}
__device__ __forceinline__ unsigned int mm_ord(float u) {
  const unsigned int b = __float_as_uint(u);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float mm_unord(unsigned int o) {
  if (o == 0u) return -CUDART_INF_F;
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
// exact ORBmatcher::DescriptorDistance of two fp32 rows, in the reference's op order (oracle/matcher_oracle.c): fp32 subtract,
// fp64 accumulate in index order, round to float, * 512, truncate
__device__ __noinline__ int mm_exact_distance_seq(const float* arow, const float* brow) {
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
// The same integer, usually 4x sooner: every square of an fp32 difference is EXACT in fp64, so any summation order differs from
// the reference's by less than 2^-45 relative (<= 67 roundings of 2^-53).  Four interleaved chains (a 16-step instead of a 64-step
// dependency chain); if rounding to float gives the same value at both ends of a 2^-40 band around the sum, that float -- hence
// the integer -- IS the reference's; otherwise (probability ~ 2^-15) the sequential order decides.
__device__ __forceinline__ int mm_exact_distance(const float* arow, const float* brow) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 8
  for (int kq = 0; kq < 16; ++kq) {
    const float4 x = *reinterpret_cast<const float4*>(arow + kq * 4);
    const float4 b = *reinterpret_cast<const float4*>(brow + kq * 4);
    float d;
    d = x.x - b.x; s0 = fma((double)d, (double)d, s0);
    d = x.y - b.y; s1 = fma((double)d, (double)d, s1);
    d = x.z - b.z; s2 = fma((double)d, (double)d, s2);
    d = x.w - b.w; s3 = fma((double)d, (double)d, s3);
  }
  const double s = (s0 + s1) + (s2 + s3);
  const float f = __double2float_rn(s);
  const float flo = __double2float_rn(s * (1.0 - 0x1p-40)), fhi = __double2float_rn(s * (1.0 + 0x1p-40));
  if (flo == fhi) return (int)(f * 512.0f);
  return mm_exact_distance_seq(arow, brow);
}

// Exact verification of a warp's queued (row, column) pairs, OUT OF LINE (the streaming loops stay small: this code runs a few
// times per CTA).  Two phases: (1) filter every entry against the current bound (rows: the row's candidate bound; columns: the
// largest estimate ANY CTA has seen for the column) and compact the survivors in place (ballot-based, warp-uniform), (2) verify
// the survivors 32 at a time, one per lane -- dense batches: the fp64 chain costs the same whether 3 or 32 lanes run it.
// ---- explicit shared-space accesses for the out-of-line drain (a generic pointer costs an address-space check per access) ----
__device__ __forceinline__ unsigned long long mm_lds64(uint32_t a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void mm_sts64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ float mm_ldsf(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ unsigned int mm_lds32(uint32_t a) { unsigned int v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ unsigned long long mm_atoms_min64(uint32_t a, unsigned long long v) {
  unsigned long long o;
  asm volatile("atom.shared.min.u64 %0, [%1], %2;" : "=l"(o) : "r"(a), "l"(v) : "memory");
  return o;
}

// One warp empties its candidate queue: (1) re-filter every entry against the bound as it stands now, compacting the queue in place;
// (2) verify the survivors exactly, FOUR LANES PER PAIR (16 dimensions each: all of a lane's 8 row loads are in flight at once, a
// round of 8 pairs costs one memory latency; one lane per pair needed 32 loads per lane, i.e. 4+ dependent batches per round:
// 9.6 k clk per round measured) and merge the exact keys.  Everything arrives in registers (the first version passed a struct by
// reference: ~10 local-memory loads per loop iteration).  flags: 1 final, 2 column direction, 4 wild (verify everything).
// tau_s: shared address of the row bounds to filter with; colg_s: shared snapshot of the column maxima or 0 (then colG, global);
// c_s: shared address of the CTA's drain constants (MmShared::dc, see MmDrainConst).
struct MmDrainConst {      // 64 bytes, written once per CTA
  const unsigned int* colG; unsigned long long* colK; const float* rawA0; const float* rawB; unsigned long long* dbg;
  float margin_col; unsigned int init_u; int row0; uint32_t k1_s; unsigned long long pad;
};
static_assert(sizeof(MmDrainConst) == 64, "drain constants");
__device__ __noinline__ void mm_drain(uint32_t q_s, int qn, int flags, int lane, uint32_t tau_s, uint32_t colg_s, uint32_t c_s) {
  const unsigned int* colG = reinterpret_cast<const unsigned int*>(mm_lds64(c_s));
  unsigned long long* colK = reinterpret_cast<unsigned long long*>(mm_lds64(c_s + 8));
  const float* rawA0 = reinterpret_cast<const float*>(mm_lds64(c_s + 16));
  const float* rawB = reinterpret_cast<const float*>(mm_lds64(c_s + 24));
  unsigned long long* dbg = reinterpret_cast<unsigned long long*>(mm_lds64(c_s + 32));
  const float margin_col = mm_ldsf(c_s + 40);
  const unsigned int init_u = mm_lds32(c_s + 44);
  const int row0 = (int)mm_lds32(c_s + 48);
  const uint32_t k1_s = mm_lds32(c_s + 52), k2_s = k1_s + 8u * MM_ROWS;
  unsigned long long* const cnt = (flags & 8) ? dbg : nullptr;        // entry counting (perturbs the timing: only on request)
  unsigned long long* const tdbg = (flags & 16) ? dbg : nullptr;      // cycle counters of one warp per direction of one CTA
  __syncwarp();
  const long long t_in = clock64();
  const bool dir2 = flags & 2, wild = flags & 4;
  int nw = 0;
  for (int base = 0; base < qn; base += 32) {
    const int i = base + lane;
    unsigned long long ent = 0;
    bool keep = false;
    if (i < qn) {
      ent = mm_lds64(q_s + 8u * (uint32_t)i);
      const uint32_t r = (uint32_t)(ent >> 24) & 0x7fu, j = (uint32_t)ent & 0xffffffu;
      const float ub = __uint_as_float((unsigned int)(ent >> 32));        // upper bound of the pair's u
      if (!dir2) keep = wild || ub > mm_ldsf(tau_s + 4u * r);
      else keep = wild || ub >= mm_unord(colg_s ? mm_lds32(colg_s + 4u * j) : __ldcg(colG + j)) - margin_col;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (keep) mm_sts64(q_s + 8u * (uint32_t)(nw + __popc(bal & ((1u << lane) - 1u))), ent);   // nw <= base: never ahead of the entries still to be read
    nw += __popc(bal);
    __syncwarp();
  }
  if (cnt && lane == 0) { atomicAdd(cnt + (dir2 ? 28 : 26), (unsigned long long)qn); atomicAdd(cnt + (dir2 ? 29 : 27), (unsigned long long)nw); }
  const long long t_f = clock64();
  const int sub = lane & 3, slot = lane >> 2;          // 4 lanes per pair, 8 pairs per round
  for (int base = 0; base < nw; base += 8) {
    const int i = base + slot;
    const bool live = i < nw;
    const unsigned long long ent = mm_lds64(q_s + 8u * (uint32_t)(live ? i : 0));
    const int r = (int)((ent >> 24) & 0x7full), j = (int)(ent & 0xffffffull);
    const float* arow = rawA0 + (size_t)r * 64 + sub * 16;
    const float* brow = rawB + (size_t)j * 64 + sub * 16;
    float4 xa[4], xb[4];
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) { xa[kq] = __ldg(reinterpret_cast<const float4*>(arow) + kq); xb[kq] = __ldg(reinterpret_cast<const float4*>(brow) + kq); }
    double acc = 0.0;
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      float d;
      d = xa[kq].x - xb[kq].x; acc = fma((double)d, (double)d, acc);
      d = xa[kq].y - xb[kq].y; acc = fma((double)d, (double)d, acc);
      d = xa[kq].z - xb[kq].z; acc = fma((double)d, (double)d, acc);
      d = xa[kq].w - xb[kq].w; acc = fma((double)d, (double)d, acc);
    }
    // every square of an fp32 difference is EXACT in fp64, so this order differs from the reference's index order by < 2^-45
    // relative; if rounding to float gives the same value at both ends of a 2^-40 band around the sum, that float -- hence the
    // integer -- IS the reference's; otherwise (probability ~ 2^-15) the sequential order decides.
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    const float f = __double2float_rn(acc);
    const float flo = __double2float_rn(acc * (1.0 - 0x1p-40)), fhi = __double2float_rn(acc * (1.0 + 0x1p-40));
    int D = (int)(f * 512.0f);
    if (live && sub == 0) {
      if (flo != fhi) D = mm_exact_distance_seq(rawA0 + (size_t)r * 64, rawB + (size_t)j * 64);
      if ((unsigned int)D < init_u) {
        if (!dir2) {
          const unsigned long long key = ((unsigned long long)(unsigned int)D << 32) | (unsigned int)j;
          const unsigned long long old = mm_atoms_min64(k1_s + 8u * (uint32_t)r, key);
          mm_atoms_min64(k2_s + 8u * (uint32_t)r, max(old, key));
        } else {
          atomicMin(colK + j, ((unsigned long long)(unsigned int)D << 32) | (unsigned int)(row0 + r));
        }
      }
    }
  }
  __syncwarp();
  if (tdbg && lane == 0) {     // slots 32.. : [dir][final] x (calls, entries in, entries kept, filter cycles, verify cycles)
    unsigned long long* d = tdbg + 32 + ((dir2 ? 2 : 0) + ((flags & 1) ? 1 : 0)) * 5;
    atomicAdd(d, 1ull); atomicAdd(d + 1, (unsigned long long)qn); atomicAdd(d + 2, (unsigned long long)nw);
    atomicAdd(d + 3, (unsigned long long)(t_f - t_in)); atomicAdd(d + 4, (unsigned long long)(clock64() - t_f));
  }
}

struct MmShared {
  uint64_t bar_a, bar_full[MM_STAGES], bar_empty[MM_STAGES], bar_accf1[2], bar_acce1[2], bar_accf2[2], bar_acce2[2];
  uint32_t tmem, flag;
  MmDrainConst dc;
};

template <bool MUTUAL>
__global__ void __launch_bounds__(MM_THREADS, 1) mm_kernel(const MatchTcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sA = smem_raw;                                                   // the CTA's 128-row block
  unsigned char* sB0 = sA + MM_BLK_BYTES;                                         // MM_STAGES column blocks
  unsigned long long* sQ1 = reinterpret_cast<unsigned long long*>(sB0 + MM_STAGES * MM_BLK_BYTES);   // [8][MM_Q1] (slice max bits << 32 | row << 24 | column)
  unsigned long long* sQ2 = sQ1 + 8 * MM_Q1;                                      // [8][MM_Q2]
  unsigned long long* sK1 = sQ2 + 8 * MM_Q2;                                      // [128] exact best key of the row
  unsigned long long* sK2 = sK1 + MM_ROWS;                                        // [128] exact second key
  float2* sM = reinterpret_cast<float2*>(sK2 + MM_ROWS);                          // [2][128] running (largest, second-largest) slice maximum per group
  float* sTau = reinterpret_cast<float*>(sM + 2 * MM_ROWS);                       // [2][128] current candidate bound per group; [0] = final after the stream
  MmShared* sh = reinterpret_cast<MmShared*>(sTau + 2 * MM_ROWS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.y;
  // XFB_MS_DEBUG: counters [20..31] = per-launch sums over CTA (0,0): cycles (total, row warp stream end / drain end, column warp stream end /
  // drain end, MMA lane loop) and, over ALL CTAs, queue entries / exact verifications per direction
  unsigned long long* dbgc = a.ms_counters;
  const bool dbg0 = dbgc != nullptr && blockIdx.x == 7 % gridDim.x && blockIdx.y == 3 % gridDim.y;   // a CTA in the middle of the launch
  const long long t_start = clock64();
  const int setA = a.pairs[2 * pair], setB = a.pairs[2 * pair + 1];
  const int nA = a.nA_dev ? min(a.nA_host, a.nA_dev[setA]) : a.nA_host;
  const int nB = a.nB_dev ? min(a.nB_host, a.nB_dev[setB]) : a.nB_host;
  const int row0 = blockIdx.x * MM_ROWS;
  const int nblk = (row0 < nA) ? (nB + MM_ROWS - 1) / MM_ROWS : 0;                // column blocks
  // The CTAs of a pair walk the column blocks in ROTATED order (CTA x starts at block x): when a CTA reaches a block, on average half
  // of the pair's other CTAs have already published their maxima of its columns, so the column direction can test against an
  // almost final bound in a single pass (see there).  kk = physical block of step k.
  const int rot = (MUTUAL && nblk > 0) ? (int)(blockIdx.x % (unsigned)nblk) : 0;
  const unsigned char* imgA = reinterpret_cast<const unsigned char*>(a.imgA) + (size_t)setA * a.img_stride_A + (size_t)blockIdx.x * MM_BLK_BYTES;
  const unsigned char* imgB = reinterpret_cast<const unsigned char*>(a.imgB) + (size_t)setB * a.img_stride_B;
  const float* rawA = a.rawA + (size_t)setA * a.raw_stride_A;
  const float* rawB = a.rawB + (size_t)setB * a.raw_stride_B;
  const unsigned int init_u = (unsigned int)a.init;
  unsigned int* colG = a.col_g + (size_t)pair * a.rows_padded_B;
  unsigned long long* colK = a.col_k + (size_t)pair * a.rows_padded_B;

  if (threadIdx.x < MM_ROWS) {
    const unsigned long long k0 = ((unsigned long long)init_u << 32) | 0xffffffffull;
    sK1[threadIdx.x] = k0; sK2[threadIdx.x] = k0;
    sTau[threadIdx.x] = -CUDART_INF_F; sTau[MM_ROWS + threadIdx.x] = -CUDART_INF_F;     // (an overflow drain before the first bound is stored verifies everything)
  }
  if (threadIdx.x == 0) {
    mbar_init(&sh->bar_a, 1);
    for (int s = 0; s < MM_STAGES; ++s) { mbar_init(&sh->bar_full[s], 1); mbar_init(&sh->bar_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->bar_accf1[s], 1); mbar_init(&sh->bar_acce1[s], 4);
      mbar_init(&sh->bar_accf2[s], 1); mbar_init(&sh->bar_acce2[s], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&sh->tmem, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem;

  // error scale of the tensor-core estimate (see header): per row for the row direction, per set for the column direction
  const float nbm = (nblk > 0) ? a.nrm_max_B[setB] : 0.f;
  const float nam = (nblk > 0) ? a.nrm_max_A[setA] : 0.f;
  const bool wild_set = !(nam < 1.0e5f && nbm < 1.0e5f);      // outside the fp16 images' range: every pair is verified

  if (warp == 0) {
    // ===== loader (one elected lane) =====
    if (elect_one_sync() && nblk > 0) {
      mbar_expect_tx(&sh->bar_a, MM_BLK_BYTES);
      bulk_g2s(sA, imgA, MM_BLK_BYTES, &sh->bar_a);
      for (int k = 0; k < nblk; ++k) {
        const int s = k % MM_STAGES;
        if (k >= MM_STAGES) mbar_wait(&sh->bar_empty[s], ((k / MM_STAGES) - 1) & 1);
        mbar_expect_tx(&sh->bar_full[s], MM_BLK_BYTES);
        const int kk = (k + rot) % nblk;
        bulk_g2s(sB0 + (size_t)s * MM_BLK_BYTES, imgB + (size_t)kk * MM_BLK_BYTES, MM_BLK_BYTES, &sh->bar_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: ONE elected lane runs the whole loop =====
    if (elect_one_sync() && nblk > 0) {
      mbar_wait(&sh->bar_a, 0);
      tc_fence_after();
      const uint32_t aAddr = smem_u32(sA);
      const uint64_t dA0 = umma_desc_kmajor(aAddr, MM_LBO, MM_SBO);
      const uint64_t dAt = umma_desc_kmajor(aAddr + 8 * MM_LBO, 2 * MM_LBO, MM_SBO);        // row tail, then the zero chunk
      constexpr uint64_t KSTEP = (2u * MM_LBO) >> 4;
#pragma unroll 1
      for (int k = 0; k < nblk; ++k) {
        const int s = k % MM_STAGES, st = k & 1;
        long long tt = dbg0 ? clock64() : 0;
        mbar_wait(&sh->bar_full[s], (k / MM_STAGES) & 1);
        if (dbg0) { const long long n2 = clock64(); atomicAdd(dbgc + 52, (unsigned long long)(n2 - tt)); tt = n2; }
        if (k >= 2) mbar_wait(&sh->bar_acce1[st], ((k >> 1) - 1) & 1);
        if (dbg0) { const long long n2 = clock64(); atomicAdd(dbgc + 54, (unsigned long long)(n2 - tt)); tt = n2; }
        tc_fence_after();
        const uint32_t bAddr = smem_u32(sB0 + (size_t)s * MM_BLK_BYTES);
        const uint64_t dB0 = umma_desc_kmajor(bAddr, MM_LBO, MM_SBO);
        const uint64_t dBt = umma_desc_kmajor(bAddr + 9 * MM_LBO, MM_LBO, MM_SBO);          // column tail, then the zero chunk
        const uint32_t d1 = tmem_base + (uint32_t)st * 256u, d2 = d1 + 128u;
        // D1[row][column] = a.b + rowtail.coltail
        mm_umma(d1, dA0, dB0, 0u);
        mm_umma(d1, dA0 + KSTEP, dB0 + KSTEP, 1u);
        mm_umma(d1, dA0 + 2 * KSTEP, dB0 + 2 * KSTEP, 1u);
        mm_umma(d1, dA0 + 3 * KSTEP, dB0 + 3 * KSTEP, 1u);
        mm_umma(d1, dAt, dBt, 1u);
        umma_commit(&sh->bar_accf1[st]);
        if (MUTUAL) {
          if (dbg0) { const long long n2 = clock64(); atomicAdd(dbgc + 55, (unsigned long long)(n2 - tt)); tt = n2; }
          if (k >= 2) { mbar_wait(&sh->bar_acce2[st], ((k >> 1) - 1) & 1); tc_fence_after(); }
          if (dbg0) { const long long n2 = clock64(); atomicAdd(dbgc + 53, (unsigned long long)(n2 - tt)); tt = n2; }
          // D2[column][row]: the same products with the operands swapped.  ORDER MATTERS: D2 is committed last, so "the column
          // direction has seen its last accumulator" implies that EVERY MMA has finished reading the ring of column blocks -- the
          // column warps reuse that ring for their snapshot of the column maxima after the stream.
          mm_umma(d2, dB0, dA0, 0u);
          mm_umma(d2, dB0 + KSTEP, dA0 + KSTEP, 1u);
          mm_umma(d2, dB0 + 2 * KSTEP, dA0 + 2 * KSTEP, 1u);
          mm_umma(d2, dB0 + 3 * KSTEP, dA0 + 3 * KSTEP, 1u);
          mm_umma(d2, dBt, dAt, 1u);
          umma_commit(&sh->bar_accf2[st]);
        }
        umma_commit(&sh->bar_empty[s]);     // the column block may be overwritten once these MMAs have read it
        if (dbg0) { const long long n2 = clock64(); atomicAdd(dbgc + 55, (unsigned long long)(n2 - tt)); }
      }
      if (dbg0) atomicAdd(dbgc + 25, (unsigned long long)(clock64() - t_start));
    }
    __syncwarp();
  } else {
    const int ew = warp - 2;                       // 0..15
    const bool dir2 = ew >= 8;                     // column direction
    const int group = (ew >> 2) & 1;               // which accumulator stage / which half of the blocks
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may access
    const int ln = quad * 32 + lane;               // line of the accumulator this lane owns
    const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)group * 256u + (dir2 ? 128u : 0u);
    unsigned long long* q = dir2 ? sQ2 + (size_t)(ew - 8) * MM_Q2 : sQ1 + (size_t)ew * MM_Q1;
    const int qcap = dir2 ? MM_Q2 : MM_Q1;
    int qn = 0;                                    // warp-uniform fill count
    const float e_col = 1.05f * sqrtf(nam * nbm) + 0.02f * (nam + nbm + 1.0f);
    const float margin_col = wild_set ? CUDART_INF_F : (2.0f * e_col + 1.0f) * (1.0f / 1024.0f);

    const uint32_t q_s = smem_u32(q), tau_s = smem_u32(sTau), c_s = smem_u32(&sh->dc);
    uint32_t colg_s = 0;                           // shared snapshot of the column maxima (column direction, final drain)
    if (lane == 0) {                               // every warp writes the same constants: no CTA barrier needed before the first overflow drain
      MmDrainConst& k = sh->dc;
      k.colG = colG; k.colK = colK; k.rawA0 = rawA + (size_t)row0 * 64; k.rawB = rawB; k.dbg = dbgc;
      k.margin_col = margin_col; k.init_u = init_u; k.row0 = row0; k.k1_s = smem_u32(sK1);
    }
    __syncwarp();
    const int dflags = (dir2 ? 2 : 0) | (wild_set ? 4 : 0) | ((a.ms_mode & 64) && dbgc ? 8 : 0) | ((dbg0 && (ew == 0 || ew == 8)) ? 16 : 0);
    auto drain = [&](bool final) {
      mm_drain(q_s, qn, dflags | (final ? 1 : 0), lane, final ? tau_s : tau_s + 4u * (uint32_t)(group * MM_ROWS), final ? colg_s : 0u, c_s);
      qn = 0;
    };
    // warp-uniform append (no atomics): every round, each lane with candidates left contributes its lowest one
    // entry e of the mask is the pair (row_base + e * row_step, col_base + e * col_step)
    auto append = [&](uint32_t mask, int row_base, int row_step, int col_base, int col_step, float ub) {
      const unsigned long long hi = (unsigned long long)__float_as_uint(ub) << 32;
      while (__any_sync(0xffffffffu, mask != 0u)) {
        if (qn > qcap - 32) drain(false);
        const bool has = mask != 0u;
        const uint32_t bal = __ballot_sync(0xffffffffu, has);
        if (has) {
          const int e = __ffs(mask) - 1;
          mask &= mask - 1u;
          const unsigned int r = (unsigned int)(row_base + e * row_step), j = (unsigned int)(col_base + e * col_step);
          q[qn + __popc(bal & ((1u << lane) - 1u))] = hi | ((unsigned long long)r << 24) | j;
        }
        qn += __popc(bal);
      }
    };

    if (!dir2) {
      // ===== row direction: lane = row =====
      const int row = row0 + ln;
      const bool ok = row < nA;
      float margin = 0.f, cap = CUDART_INF_F;
      bool wild = wild_set;
      if (ok && nblk > 0) {
        const float na = a.nrmA[(size_t)setA * a.rows_padded_A + row];
        const float e = 1.05f * sqrtf(na * nbm) + 0.02f * (na + nbm + 1.0f);     // |t - 512 float(d)| <= e
        margin = (2.0f * e + 1.0f) * (1.0f / 1024.0f);
        cap = (a.init == 0x7fffffff) ? -CUDART_INF_F : -((float)a.init + e) * (1.0f / 1024.0f);   // t < init + e  <=>  u > cap
        wild = wild || !(na < 1.0e5f);
      }
      const bool need2 = a.second_dist != nullptr;
      float m1 = -CUDART_INF_F, m2 = -CUDART_INF_F;
      auto bound = [&]() {
        if (!ok) return CUDART_INF_F;
        if (wild) return -CUDART_INF_F;
        return fmaxf((need2 ? m2 : m1) - margin, cap);
      };
#pragma unroll 1
      for (int k = group; k < nblk; k += 2) {
        const bool tme = dbg0 && ew == 0 && lane == 0;
        long long tt = tme ? clock64() : 0;
        mbar_wait(&sh->bar_accf1[group], (uint32_t)(k >> 1) & 1u);
        __syncwarp();
        tc_fence_after();
        if (tme) { const long long n2 = clock64(); atomicAdd(dbgc + 56, (unsigned long long)(n2 - tt)); tt = n2; }
        float v[32];
        if (k == group) {
          // seed the running top-2 from this block before collecting candidates from it
#pragma unroll 1
          for (int part = 0; part < 4; ++part) {
            mm_ld32(tq + (uint32_t)part * 32u, v);
            const float m = mm_max32(v);
            const float lo = fminf(m1, m);
            m1 = fmaxf(m1, m);
            m2 = fmaxf(m2, lo);
          }
        }
#pragma unroll 1
        for (int part = 0; part < 4; ++part) {
          const float tau = bound();
          mm_ld32(tq + (uint32_t)part * 32u, v);
          const float m = mm_max32(v);
          // once the running bound is tight most slices hold no candidate for ANY of the warp's 32 rows (the slice maximum says so)
          uint32_t mask = 0u;
          if (__any_sync(0xffffffffu, m > tau)) mask = mm_mask32(v, tau);
          if (wild && ok) mask = 0xffffffffu;
          const int j0 = ((k + rot) % nblk) * MM_ROWS + part * 32;
          if (j0 + 32 > nB) mask &= (nB > j0) ? (0xffffffffu >> (32 - (nB - j0))) : 0u;     // padded columns
          if (k != group) {            // (the seeding block's maxima are already in)
            const float lo = fminf(m1, m);
            m1 = fmaxf(m1, m);
            m2 = fmaxf(m2, lo);
          }
          append(mask, ln, 0, j0, 1, m);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->bar_acce1[group]);
        sTau[group * MM_ROWS + ln] = bound();          // what an overflow drain of this group filters with
        if (tme) atomicAdd(dbgc + 57, (unsigned long long)(clock64() - tt));
      }
      // ---- the stream is over: merge the two groups' slice maxima into the final bound of every row ----
      sM[group * MM_ROWS + ln] = make_float2(m1, m2);
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (group == 0) {
        const float2 o = sM[MM_ROWS + ln];
        const float M1 = fmaxf(m1, o.x), M2 = fmaxf(fminf(m1, o.x), fmaxf(m2, o.y));
        m1 = M1; m2 = M2;
        sTau[ln] = bound();
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (dbg0 && ew == 0 && lane == 0) atomicAdd(dbgc + 21, (unsigned long long)(clock64() - t_start));
      drain(true);
      if (dbg0 && ew == 0 && lane == 0) atomicAdd(dbgc + 22, (unsigned long long)(clock64() - t_start));
    } else if (MUTUAL) {
      // ===== column direction: lane = column of the block.  ONE pass over the accumulator: a row is a candidate when its value
      // exceeds (the largest value known for the column) - margin, where "known" = the maximum the pair's other CTAs have published
      // so far (prefetched one block ahead; the rotated block order makes it nearly final on average) and this CTA's own rows up to
      // and including the current 32-row part.  Any such bound is <= the final column maximum, so no true candidate is lost; the
      // final filter of the drain compares with the final maximum.  Most parts then hold no candidate for any of the warp's 32
      // columns and skip the mask altogether.
      const int rows_ok = min(MM_ROWS, nA - row0);
      unsigned int gnext = 0u;
      if (group < nblk) { const int j1 = ((group + rot) % nblk) * MM_ROWS + ln; if (j1 < nB) gnext = __ldcg(colG + j1); }
#pragma unroll 1
      for (int k = group; k < nblk; k += 2) {
        const int j = ((k + rot) % nblk) * MM_ROWS + ln;
        const bool colok = j < nB;
        float run = mm_unord(gnext);                      // (-inf when nobody has published yet)
        gnext = 0u;
        if (k + 2 < nblk) { const int j2 = ((k + 2 + rot) % nblk) * MM_ROWS + ln; if (j2 < nB) gnext = __ldcg(colG + j2); }
        const bool tme = dbg0 && ew == 8 && lane == 0;
        long long tt = tme ? clock64() : 0;
        mbar_wait(&sh->bar_accf2[group], (uint32_t)(k >> 1) & 1u);
        __syncwarp();
        tc_fence_after();
        if (tme) { const long long n2 = clock64(); atomicAdd(dbgc + 58, (unsigned long long)(n2 - tt)); tt = n2; }
        float v[32];
        float own = -CUDART_INF_F;
#pragma unroll 1
        for (int part = 0; part < 4; ++part) {
          mm_ld32(tq + (uint32_t)part * 32u, v);
          const float m = mm_max32(v);
          own = fmaxf(own, m);
          run = fmaxf(run, m);
          const float thr = run - margin_col;
          uint32_t mask = 0u;
          if (__any_sync(0xffffffffu, m > thr)) mask = mm_mask32(v, thr);
          if (wild_set) mask = 0xffffffffu;
          const int r0 = part * 32;
          if (r0 + 32 > rows_ok) mask &= (rows_ok > r0) ? (0xffffffffu >> (32 - (rows_ok - r0))) : 0u;   // padded rows
          if (!colok) mask = 0u;
          append(mask, r0, 1, j, 0, m);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->bar_acce2[group]);
        if (colok) atomicMax(colG + j, mm_ord(own));
        if (tme) atomicAdd(dbgc + 59, (unsigned long long)(clock64() - tt));
      }
      if (dbg0 && ew == 8 && lane == 0) atomicAdd(dbgc + 23, (unsigned long long)(clock64() - t_start));
      // The final filter compares every queued row with the pair's per-column maximum over ALL CTAs so far: a dependent L2 load per
      // 32 entries (27 k clk for ~550 entries, measured).  The stream is over -- both groups are past their last accumulator, so
      // every MMA has read its column block -- and the ring of column blocks is free: park a snapshot of the maxima there (any
      // older value is a valid, merely looser, bound).
      asm volatile("bar.sync 3, 256;" ::: "memory");
      if (a.rows_padded_B * 4 <= MM_STAGES * MM_BLK_BYTES) {
        unsigned int* sG = reinterpret_cast<unsigned int*>(sB0);
        for (int j = (ew - 8) * 32 + lane; j < a.rows_padded_B; j += 256) sG[j] = __ldcg(colG + j);
        colg_s = smem_u32(sG);
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");
      drain(true);
      if (dbg0 && ew == 8 && lane == 0) atomicAdd(dbgc + 24, (unsigned long long)(clock64() - t_start));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (dbg0 && threadIdx.x == 0) { atomicAdd(dbgc + 20, (unsigned long long)(clock64() - t_start)); atomicAdd(dbgc + 30, 1ull); }
  if (threadIdx.x < MM_ROWS) {
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
  if (MUTUAL) {
    // the last CTA of the pair turns the merged column keys into the outputs and resets the scratch for the next launch
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(a.pair_done + pair, 1u);
      sh->flag = (prev == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (sh->flag) {
      __threadfence();
      for (int j = threadIdx.x; j < a.rows_padded_B; j += MM_THREADS) {
        const unsigned long long key = __ldcg(colK + j);
        if (j < a.out_stride_cols) {
          const bool have = j < nB && key != ~0ull;
          const size_t o = (size_t)pair * a.out_stride_cols + j;
          if (a.rev_idx) a.rev_idx[o] = have ? (int)(unsigned int)(key & 0xffffffffull) : -1;
          if (a.rev_dist) a.rev_dist[o] = have ? (int)(key >> 32) : a.init;
        }
        colK[j] = ~0ull;
        colG[j] = 0u;
      }
      if (threadIdx.x == 0) a.pair_done[pair] = 0u;
    }
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

constexpr size_t MM_SMEM = (size_t)(1 + MM_STAGES) * MM_BLK_BYTES + (size_t)8 * (MM_Q1 + MM_Q2) * 8 + (size_t)MM_ROWS * (8 + 8 + 2 * 8 + 2 * 4) +
                           sizeof(MmShared) + 64;
static_assert(MM_SMEM <= 227 * 1024, "shared memory budget");

cudaError_t launch_mm_prep(Ctx* c, const float* desc, size_t set_stride, int n_sets, const int32_t* n_dev, int n_host, int rows_padded,
                           void* img, size_t img_set_bytes, float* nrm, float* nrm_max) {
  dim3 grid(rows_padded / MM_ROWS, n_sets);
  cudaError_t e0 = cudaMemsetAsync(nrm_max, 0, (size_t)n_sets * 4, c->stream);
  if (e0 != cudaSuccess) return e0;
  prof_begin(c, P_MATCH_PREP);
  mm_prep_kernel<<<grid, 256, 0, c->stream>>>(desc, set_stride, n_dev, n_host, rows_padded, reinterpret_cast<unsigned char*>(img), img_set_bytes, nrm, nrm_max);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

// a.img_stride_* in BYTES; a.col_g / a.col_k / a.pair_done: per-pair column scratch (zero / all-ones / zero between launches)
cudaError_t launch_match_mutual(Ctx* c, const MatchTcArgs& a, int n_pairs, bool mutual) {
  static unsigned long long attr_mask = 0;
  if (!((attr_mask >> c->device) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(mm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MM_SMEM);
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << c->device;
  }
  prof_begin(c, P_MATCH_TILE);
  if (mutual) mm_kernel<true><<<dim3(a.rows_padded_A / MM_ROWS, n_pairs), MM_THREADS, MM_SMEM, c->stream>>>(a);
  else mm_kernel<false><<<dim3(a.rows_padded_A / MM_ROWS, n_pairs), MM_THREADS, MM_SMEM, c->stream>>>(a);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

}  // namespace xfb
