// post.cu -- discrete post-processing of XFextractor::operator() (src/XFextractor.cc:273-316):
//   nms_score : 5x5 max-pool NMS (XFextractor::NMS, :219-248: x == local_max && x > thr) fused with the
//               reliability score nearest(K1h)(kp) * bilinear(H1)(kp) (:280) and the (0,0) mask (:281-282);
//               every surviving candidate is appended as one sortable 64-bit key
//   topk      : per-frame exact top-k: bitonic sort of all candidates when they fit shared memory, else MSB radix passes first (:285-295).  The
//               reference's argsort is unstable; this build's order is score descending, then row-major
//               pixel index ascending (SURVEY.md hard part (c)) -- the key encodes exactly that.
//   describe  : InterpolateSparse2d bilinear on the channel-normalised feature map, then L2-normalise
//               (:273, :298-301); one warp per keypoint.
// grid_sample geometry follows ATen's vectorised CPU kernel (the path the reference takes for fp32):
//   g = 2*(p/(S-1)) - 1 (InterpolateSparse2d::normgrid, src/XFeat.cc:181-186),
//   src = (g + 1) * (S_map / 2) - 0.5, zero padding, align_corners = false.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "xfb_internal.h"

namespace xfb {

typedef unsigned long long u64;

// Rounding sequence pinned against ATen's AVX grid_sampler (tools/pin_gridsample.py): the
// un-normalisation is ONE fma, (g + 1) * (S_map / 2) - 0.5; explicit _rn intrinsics keep nvcc from
// re-associating or contracting anything else.
__device__ __forceinline__ float grid_src(int p, int full, int map) {
  const float q = __fdiv_rn((float)p, (float)(full - 1));
  const float g = __fsub_rn(__fmul_rn(2.0f, q), 1.0f);
  return __fmaf_rn(__fadd_rn(g, 1.0f), __fdiv_rn((float)map, 2.0f), -0.5f);
}
// out = fma(se_v, se, fma(sw_v, sw, fma(ne_v, ne, nw_v * nw)))  -- the contraction ATen's kernel compiles to
__device__ __forceinline__ float bilerp4(float nwv, float nev, float swv, float sev, float nw, float ne, float sw, float se) {
  return __fmaf_rn(sev, se, __fmaf_rn(swv, sw, __fmaf_rn(nev, ne, __fmul_rn(nwv, nw))));
}

// bilinear sample of a single-channel map with zero padding
__device__ __forceinline__ float bilinear1(const float* map, int mh, int mw, float sx, float sy) {
  const float x0f = floorf(sx), y0f = floorf(sy);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const float w = __fsub_rn(sx, x0f), e = __fsub_rn(1.0f, w), n = __fsub_rn(sy, y0f), s = __fsub_rn(1.0f, n);
  const bool xin0 = x0 >= 0 && x0 < mw, xin1 = x0 + 1 >= 0 && x0 + 1 < mw;
  const bool yin0 = y0 >= 0 && y0 < mh, yin1 = y0 + 1 >= 0 && y0 + 1 < mh;
  const float nw = (xin0 && yin0) ? map[(size_t)y0 * mw + x0] : 0.f;
  const float ne = (xin1 && yin0) ? map[(size_t)y0 * mw + x0 + 1] : 0.f;
  const float sw = (xin0 && yin1) ? map[(size_t)(y0 + 1) * mw + x0] : 0.f;
  const float se = (xin1 && yin1) ? map[(size_t)(y0 + 1) * mw + x0 + 1] : 0.f;
  return bilerp4(nw, ne, sw, se, __fmul_rn(s, e), __fmul_rn(s, w), __fmul_rn(n, e), __fmul_rn(n, w));
}

constexpr int NMS_T = 16;    // 256-thread CTAs (16 x 16) ...
constexpr int NMS_NX = 4;    // ... that each sweep a 64 x 16 pixel strip: 4x fewer, 4x longer CTAs than one 16 x 16 tile each
                             // (the kernel is launch- / tile-load-latency bound, not bandwidth bound)
__global__ void __launch_bounds__(NMS_T * NMS_T) nms_score_kernel(const float* k1h, const float* h1, int H, int W, float thr,
                                                                  u64* cand, int* cand_count) {
  constexpr int TW = NMS_T * NMS_NX;
  __shared__ float tile[NMS_T + 4][TW + 4];
  const int b = blockIdx.z;
  const float* img = k1h + (size_t)b * H * W;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * NMS_T;
  const int tid = threadIdx.y * NMS_T + threadIdx.x;
  for (int i = tid; i < (NMS_T + 4) * (TW + 4); i += NMS_T * NMS_T) {
    const int ty = i / (TW + 4), tx = i - ty * (TW + 4);
    const int gy = y0 + ty - 2, gx = x0 + tx - 2;
    tile[ty][tx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? img[(size_t)gy * W + gx] : -CUDART_INF_F;
  }
  __syncthreads();
  __shared__ int s_cnt[NMS_NX * 8];   // candidates per (sub-tile, warp), then their exclusive prefix
  __shared__ int s_base;
  const int lane = tid & 31, wrp = tid >> 5;
  u64 keys[NMS_NX];
  unsigned int ballots[NMS_NX];
#pragma unroll
  for (int sx = 0; sx < NMS_NX; ++sx) {
    const int lx = sx * NMS_T + threadIdx.x;   // column inside the strip
    const int x = x0 + lx, y = y0 + threadIdx.y;
    u64 key = 0;   // 0 = not a candidate (a real key has score bits > 0)
    if (x < W && y < H) {
      const float v = tile[threadIdx.y + 2][lx + 2];
      if (v > thr) {
        float m = v;
#pragma unroll
        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) m = fmaxf(m, tile[threadIdx.y + dy][lx + dx]);
        if (v == m && !(x == 0 && y == 0)) {   // (0,0) is masked to -1 by the reference, never valid
          // nearest(K1h)(kp): grid_sample nearest at full resolution (drops the last row / column)
          const float nx = nearbyintf(grid_src(x, W, W)), ny = nearbyintf(grid_src(y, H, H));
          float sn = 0.f;
          if (nx >= 0.f && nx < (float)W && ny >= 0.f && ny < (float)H) sn = img[(size_t)(int)ny * W + (int)nx];
          const int mh = H >> 3, mw = W >> 3;
          const float sb = bilinear1(h1 + (size_t)b * mh * mw, mh, mw, grid_src(x, W, mw), grid_src(y, H, mh));
          const float score = __fmul_rn(sn, sb);
          if (score > 0.f) {                   // `valid = scores > 0`, src/XFextractor.cc:313
            const unsigned int lin = (unsigned int)(y * W + x);
            key = ((u64)__float_as_uint(score) << 32) | (u64)(0xFFFFFFFFu - lin);
          }
        }
      }
    }
    keys[sx] = key;
    ballots[sx] = __ballot_sync(0xffffffffu, key != 0);
    if (lane == 0) s_cnt[sx * 8 + wrp] = __popc(ballots[sx]);
  }
  __syncthreads();
  // CTA-aggregated append: ONE atomic per CTA on the frame's counter (the 32 per-frame counters share a cache line, so
  // per-warp atomics of the whole batch serialise on it); the order inside the list is irrelevant, the key sorts
  if (tid == 0) {
    int total = 0;
    for (int i = 0; i < NMS_NX * 8; ++i) { const int n = s_cnt[i]; s_cnt[i] = total; total += n; }
    s_base = total ? atomicAdd(cand_count + b * XFB_TICKET_STRIDE, total) : 0;
  }
  __syncthreads();
#pragma unroll
  for (int sx = 0; sx < NMS_NX; ++sx)
    if (keys[sx] != 0) cand[(size_t)b * H * W + s_base + s_cnt[sx * 8 + wrp] + __popc(ballots[sx] & ((1u << lane) - 1u))] = keys[sx];
}

// ------------------------------------------------------------------------------------------------
constexpr int TOPK_NT = 1024;
constexpr int TOPK_CAP = 8192;     // keys the shared-memory sorter holds
// Exact per-frame top-k of the candidate keys (score bits << 32 | ~pixel index: unique, so "the k largest keys" is the reference's
// order with this build's tie-break).  If the frame has at most TOPK_CAP candidates they are all sorted (bitonic, shared memory)
// and the first k taken.  Otherwise MSB-first radix passes (one private 256-bin histogram per warp: no inter-warp contention on
// the few hot exponent bins) narrow the selection until the keys >= the running prefix fit the sorter -- usually two passes.
__global__ void __launch_bounds__(TOPK_NT) topk_kernel(const u64* cand, int* cand_count, int* cand_count_last, int HW, int W,
                                                       int topk, int32_t* n_valid, float* kpt_xy, float* score) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* skeys = reinterpret_cast<u64*>(smem_raw);                                  // [TOPK_CAP]
  unsigned int* whist = reinterpret_cast<unsigned int*>(skeys + TOPK_CAP);        // [32 warps][256]
  __shared__ unsigned int hist[256];
  __shared__ u64 s_prefix;
  __shared__ int s_remaining, s_take, s_done, s_fill;
  const int b = blockIdx.x, t = threadIdx.x, wrp = t >> 5;
  const u64* keys = cand + (size_t)b * HW;
  const int N = cand_count[b * XFB_TICKET_STRIDE];
  const int K = N < topk ? N : topk;
  u64 thresh = 0;      // select keys >= thresh
  int take = N;        // how many keys that is
  if (N > TOPK_CAP) {
    if (t == 0) { s_prefix = 0; s_remaining = K; s_done = 0; }
    u64 mask = 0;
    for (int pass = 7; pass >= 0; --pass) {
      for (int i = t; i < 32 * 256; i += TOPK_NT) whist[i] = 0;
      __syncthreads();
      const u64 prefix = s_prefix;
      const int shift = pass * 8;
      for (int i = t; i < N; i += TOPK_NT) {
        const u64 k = keys[i];
        if ((k & mask) == prefix) atomicAdd(&whist[wrp * 256 + ((unsigned int)(k >> shift) & 255u)], 1u);
      }
      __syncthreads();
      if (t < 256) {
        unsigned int h = 0;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) h += whist[w * 256 + t];
        hist[t] = h;
      }
      __syncthreads();
      if (t == 0) {
        int rem = s_remaining;
        int d = 255;
        for (; d > 0; --d) {
          const int hcount = (int)hist[d];
          if (hcount >= rem) break;
          rem -= hcount;
        }
        s_prefix = prefix | ((u64)d << shift);
        s_remaining = rem;
        s_take = K - rem + (int)hist[d];            // keys >= the new prefix: everything already selected + the whole boundary bin
        s_done = (s_take <= TOPK_CAP) ? 1 : 0;
      }
      mask |= (u64)255 << shift;
      __syncthreads();
      if (s_done) break;
    }
    thresh = s_prefix;   // (after 8 passes: the K-th largest key itself, take == K)
    take = s_take;
  }
  int sort_n = 2;
  while (sort_n < take) sort_n <<= 1;
  if (t == 0) s_fill = 0;
  for (int i = t; i < sort_n; i += TOPK_NT) skeys[i] = 0;
  __syncthreads();
  // compaction, warp-aggregated: ONE shared-memory atomic per warp and round (a per-key atomic on one address serialises the CTA)
  for (int i0 = 0; i0 < N; i0 += TOPK_NT) {
    const int i = i0 + t;
    const u64 k = (i < N) ? keys[i] : 0;
    const bool sel = i < N && k >= thresh;
    const unsigned int bal = __ballot_sync(0xffffffffu, sel);
    int base = 0;
    if ((t & 31) == 0 && bal) base = atomicAdd(&s_fill, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (sel) {
      const int pos = base + __popc(bal & ((1u << (t & 31)) - 1u));
      if (pos < sort_n) skeys[pos] = k;
    }
  }
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= sort_n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = t; i < (sort_n >> 1); i += TOPK_NT) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const u64 a = skeys[lo], c = skeys[hi];
        if ((a < c) == desc) { skeys[lo] = c; skeys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = t; i < topk; i += TOPK_NT) {
    float sc = 0.f, fx = 0.f, fy = 0.f;
    if (i < K) {
      const u64 k = skeys[i];
      sc = __uint_as_float((unsigned int)(k >> 32));
      const unsigned int lin = 0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFu);
      fx = (float)(lin % (unsigned int)W);
      fy = (float)(lin / (unsigned int)W);
    }
    score[(size_t)b * topk + i] = sc;
    kpt_xy[((size_t)b * topk + i) * 2] = fx;
    kpt_xy[((size_t)b * topk + i) * 2 + 1] = fy;
  }
  if (t == 0) {
    n_valid[b] = K;
    cand_count_last[b] = N;
    cand_count[b * XFB_TICKET_STRIDE] = 0;  // re-arm for the next call
  }
}

// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// channel-normalised corner: feats[y][x][:] / max(||.||_2, 1e-12)  (F.normalize(M1, dim=1), :273)
__device__ __forceinline__ float2 corner_unit(const float* fmap, int mh, int mw, int x, int y, int lane) {
  if (x < 0 || x >= mw || y < 0 || y >= mh) return make_float2(0.f, 0.f);
  const float2 v = *reinterpret_cast<const float2*>(fmap + ((size_t)y * mw + x) * 64 + lane * 2);
  const float nrm = sqrtf(warp_sum(v.x * v.x + v.y * v.y));
  const float den = fmaxf(nrm, 1e-12f);
  return make_float2(__fdiv_rn(v.x, den), __fdiv_rn(v.y, den));
}

__global__ void __launch_bounds__(256) describe_kernel(const float* feats, int H, int W, int topk, const int32_t* n_valid,
                                                       const float* kpt_xy, float* desc) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= topk) return;
  float2 o = make_float2(0.f, 0.f);
  if (i < n_valid[b]) {
    const int mh = H >> 3, mw = W >> 3;
    const float* fmap = feats + (size_t)b * mh * mw * 64;
    const int px = (int)kpt_xy[((size_t)b * topk + i) * 2], py = (int)kpt_xy[((size_t)b * topk + i) * 2 + 1];
    const float sx = grid_src(px, W, mw), sy = grid_src(py, H, mh);
    const float x0f = floorf(sx), y0f = floorf(sy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float w = __fsub_rn(sx, x0f), e = __fsub_rn(1.0f, w), n = __fsub_rn(sy, y0f), s = __fsub_rn(1.0f, n);
    const float2 nw = corner_unit(fmap, mh, mw, x0, y0, lane);
    const float2 ne = corner_unit(fmap, mh, mw, x0 + 1, y0, lane);
    const float2 sw = corner_unit(fmap, mh, mw, x0, y0 + 1, lane);
    const float2 se = corner_unit(fmap, mh, mw, x0 + 1, y0 + 1, lane);
    const float wnw = __fmul_rn(s, e), wne = __fmul_rn(s, w), wsw = __fmul_rn(n, e), wse = __fmul_rn(n, w);
    float2 v;
    v.x = bilerp4(nw.x, ne.x, sw.x, se.x, wnw, wne, wsw, wse);
    v.y = bilerp4(nw.y, ne.y, sw.y, se.y, wnw, wne, wsw, wse);
    const float den = fmaxf(sqrtf(warp_sum(v.x * v.x + v.y * v.y)), 1e-12f);
    o = make_float2(__fdiv_rn(v.x, den), __fdiv_rn(v.y, den));
  }
  *reinterpret_cast<float2*>(desc + ((size_t)b * topk + i) * 64 + lane * 2) = o;
}

cudaError_t launch_post(Ctx* c, int topk, float nms_thr, int32_t* d_nvalid, float* d_xy, float* d_score, float* d_desc) {
  const int H = c->H, W = c->W;
  dim3 g1((W + NMS_T * NMS_NX - 1) / (NMS_T * NMS_NX), (H + NMS_T - 1) / NMS_T, c->B);
  prof_begin(c, P_NMS);
  nms_score_kernel<<<g1, dim3(NMS_T, NMS_T), 0, c->stream>>>(c->k1h, c->act[L_HM_2], H, W, nms_thr, c->cand, c->cand_count);
  prof_end(c);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t smem = sizeof(u64) * (size_t)TOPK_CAP + 32 * 256 * 4;
  static unsigned long long attr_mask = 0;
  if (!((attr_mask >> c->device) & 1ull)) {
    e = cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_mask |= 1ull << c->device;
  }
  prof_begin(c, P_TOPK);
  topk_kernel<<<c->B, TOPK_NT, smem, c->stream>>>(c->cand, c->cand_count, c->cand_count_last, H * W, W, topk, d_nvalid, d_xy, d_score);
  prof_end(c);
  c->launches++;
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 g3((topk + 7) / 8, c->B);
  prof_begin(c, P_DESCRIBE);
  describe_kernel<<<g3, 256, 0, c->stream>>>(c->act[L_F_2], H, W, topk, d_nvalid, d_xy, d_desc);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

}  // namespace xfb
