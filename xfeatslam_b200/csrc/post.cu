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

// 64 x 16 output pixels per 256-thread CTA.  Three phases, each with every lane busy:
//   1. the (64 + 8) x (16 + 4) halo tile comes in as aligned float4 (x0 - 4 .. x0 + 67);
//   2. SEPARABLE 5x5 maximum: a thread owns 4 consecutive columns -- horizontal 5-maxima of all 20 rows into shared memory (3 LDS.128
//      + 9 FMNMX per 4 values), then the vertical 5-maximum of its 4 output pixels (6 LDS.128); a pixel with x == max && x > thr goes
//      onto a shared-memory list (about one pixel in 30);
//   3. the reliability score (two IEEE divisions per coordinate, a bilinear tap: ~300 instructions) is computed for the LISTED pixels
//      only, one per thread -- the round-1 kernel ran that path divergently in every warp that held a maximum (70 M warp
//      instructions per 32 VGA frames, 92 us; ncu source view, profiles/r02_step_ncu_full_table.md).
// The CTA appends its keys with ONE atomic on the frame's counter; the order inside the list is irrelevant, the key sorts.
constexpr int NMS_TW = 64, NMS_TH = 16, NMS_NT = 256;
constexpr int NMS_IW = NMS_TW + 8, NMS_IH = NMS_TH + 4;
__global__ void __launch_bounds__(NMS_NT) nms_score_kernel(const float* k1h, const float* h1, int H, int W, float thr,
                                                           u64* cand, int* cand_count) {
  __shared__ __align__(16) float sbuf[NMS_IH * NMS_IW + NMS_IH * NMS_TW];
  float (*tin)[NMS_IW] = reinterpret_cast<float (*)[NMS_IW]>(sbuf);
  float (*hmx)[NMS_TW] = reinterpret_cast<float (*)[NMS_TW]>(sbuf + NMS_IH * NMS_IW);
  __shared__ unsigned short s_list[NMS_TW * NMS_TH];
  __shared__ int s_n, s_nkeys, s_base;
  u64* s_keys = reinterpret_cast<u64*>(sbuf);                // [<= 1024] reuses the two tiles (5760 + 5120 B >= 8192 B) after phase 2
  static_assert(sizeof(float) * NMS_IH * NMS_IW % 16 == 0 && sizeof(sbuf) >= 8 * NMS_TW * NMS_TH, "key buffer");
  const int b = blockIdx.z, t = threadIdx.x;
  const float* img = k1h + (size_t)b * H * W;
  const int x0 = blockIdx.x * NMS_TW, y0 = blockIdx.y * NMS_TH;
  if (t == 0) { s_n = 0; s_nkeys = 0; }
  // ---- 1. halo tile ----
  const bool vec = (W & 3) == 0;
  for (int i = t; i < NMS_IH * (NMS_IW / 4); i += NMS_NT) {
    const int r = i / (NMS_IW / 4), c4 = i - r * (NMS_IW / 4);
    const int gy = y0 - 2 + r, gx = x0 - 4 + 4 * c4;
    float4 v = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    if (gy >= 0 && gy < H) {
      if (vec && gx >= 0 && gx + 3 < W) v = *reinterpret_cast<const float4*>(img + (size_t)gy * W + gx);
      else {
        const float* row = img + (size_t)gy * W;
        if (gx >= 0 && gx < W) v.x = row[gx];
        if (gx + 1 >= 0 && gx + 1 < W) v.y = row[gx + 1];
        if (gx + 2 >= 0 && gx + 2 < W) v.z = row[gx + 2];
        if (gx + 3 >= 0 && gx + 3 < W) v.w = row[gx + 3];
      }
    }
    *reinterpret_cast<float4*>(&tin[r][4 * c4]) = v;
  }
  __syncthreads();
  // ---- 2a. horizontal 5-maxima: hmx[r][x] = max tin[r][x + 2 .. x + 6]  (tile column = x + 4) ----
  for (int i = t; i < NMS_IH * (NMS_TW / 4); i += NMS_NT) {
    const int r = i / (NMS_TW / 4), xg = i - r * (NMS_TW / 4);
    const float4 p = *reinterpret_cast<const float4*>(&tin[r][4 * xg]);
    const float4 q = *reinterpret_cast<const float4*>(&tin[r][4 * xg + 4]);
    const float4 u = *reinterpret_cast<const float4*>(&tin[r][4 * xg + 8]);
    // columns 4xg+2 .. 4xg+9 = p.z p.w q.x q.y q.z q.w u.x u.y
    const float c = fmaxf(fmaxf(q.x, q.y), q.z);             // shared by outputs 0 and 1 (+ q.w: by 1 and 2 ...)
    float4 o;
    o.x = fmaxf(fmaxf(p.z, p.w), c);
    o.y = fmaxf(fmaxf(p.w, q.w), c);
    const float d = fmaxf(fmaxf(q.y, q.z), q.w);
    o.z = fmaxf(fmaxf(q.x, u.x), d);
    o.w = fmaxf(fmaxf(u.x, u.y), d);
    *reinterpret_cast<float4*>(&hmx[r][4 * xg]) = o;
  }
  __syncthreads();
  // ---- 2b. vertical 5-maximum of this thread's 4 pixels, candidate test ----
  {
    const int ty = t >> 4, xg = t & 15;
    const float4 v = *reinterpret_cast<const float4*>(&tin[ty + 2][4 * xg + 4]);
    float4 m = *reinterpret_cast<const float4*>(&hmx[ty][4 * xg]);
#pragma unroll
    for (int dy = 1; dy < 5; ++dy) {
      const float4 h = *reinterpret_cast<const float4*>(&hmx[ty + dy][4 * xg]);
      m.x = fmaxf(m.x, h.x); m.y = fmaxf(m.y, h.y); m.z = fmaxf(m.z, h.z); m.w = fmaxf(m.w, h.w);
    }
    const int y = y0 + ty, xb = x0 + 4 * xg;
    const float vv[4] = {v.x, v.y, v.z, v.w}, mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = xb + j;
      // x == local_max && x > thr (:236-240); (0,0) is masked to -1 by the reference, never valid
      if (vv[j] > thr && vv[j] == mm[j] && x < W && y < H && !(x == 0 && y == 0)) s_list[atomicAdd(&s_n, 1)] = (unsigned short)((ty << 6) | (4 * xg + j));
    }
  }
  __syncthreads();
  // ---- 3. scores of the listed pixels, one per thread ----
  const int n = s_n;
  for (int i0 = 0; i0 < n; i0 += NMS_NT) {          // (n <= 256 except on plateaus)
    const int i = i0 + t;
    u64 key = 0;   // 0 = not a candidate (a real key has score bits > 0)
    if (i < n) {
      const int e = s_list[i];
      const int x = x0 + (e & 63), y = y0 + (e >> 6);
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
    // (the first pass may overwrite tin / hmx: every thread is past phase 2)
    const unsigned int bal = __ballot_sync(0xffffffffu, key != 0);
    int wbase = 0;
    if ((t & 31) == 0 && bal) wbase = atomicAdd(&s_nkeys, __popc(bal));
    wbase = __shfl_sync(0xffffffffu, wbase, 0);
    if (key != 0) s_keys[wbase + __popc(bal & ((1u << (t & 31)) - 1u))] = key;
  }
  __syncthreads();
  const int nk = s_nkeys;
  if (nk == 0) return;
  if (t == 0) s_base = atomicAdd(cand_count + b * XFB_TICKET_STRIDE, nk);
  __syncthreads();
  u64* dst = cand + (size_t)b * H * W + s_base;
  for (int i = t; i < nk; i += NMS_NT) dst[i] = s_keys[i];
}

// ------------------------------------------------------------------------------------------------
// ---- bitonic network pieces for the register-resident sorter (thread t holds elements 8t .. 8t+7) ----
__device__ __forceinline__ void tk_cx(u64& a, u64& b, bool desc) {      // afterwards a >= b when desc, a <= b otherwise
  const bool sw = (a < b) == desc;
  const u64 x = sw ? b : a;
  b = sw ? a : b;
  a = x;
}
__device__ __forceinline__ void tk_local(u64 (&k)[8], bool desc) {      // strides 4, 2, 1 inside the thread
  tk_cx(k[0], k[4], desc); tk_cx(k[1], k[5], desc); tk_cx(k[2], k[6], desc); tk_cx(k[3], k[7], desc);
  tk_cx(k[0], k[2], desc); tk_cx(k[1], k[3], desc); tk_cx(k[4], k[6], desc); tk_cx(k[5], k[7], desc);
  tk_cx(k[0], k[1], desc); tk_cx(k[2], k[3], desc); tk_cx(k[4], k[5], desc); tk_cx(k[6], k[7], desc);
}
__device__ __forceinline__ void tk_keep(u64 (&k)[8], const u64 (&o)[8], bool keepmax) {
#pragma unroll
  for (int r = 0; r < 8; ++r) k[r] = keepmax ? (k[r] > o[r] ? k[r] : o[r]) : (k[r] < o[r] ? k[r] : o[r]);
}
// exchange layout in shared memory: thread t's 16-byte chunk c sits at chunk slot c ^ ((t >> 1) & 3) of its 64-byte block, so the
// 8 lanes of a quarter warp touch 8 distinct 16-byte bank groups (a plain 64-byte stride would be a 4-way conflict)
__device__ __forceinline__ ulonglong2* tk_slot(u64* skeys, int t, int c) { return reinterpret_cast<ulonglong2*>(skeys + 8 * t + 2 * (c ^ ((t >> 1) & 3))); }

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
  if (sort_n > TOPK_NT) {
    // ---- register-resident bitonic sort of TOPK_CAP keys (zero padded), descending: 8 keys per thread.  Strides 1, 2, 4 are
    // compare-exchanges inside the thread, strides 8 .. 128 one 64-bit shuffle per key, and only strides >= 256 (15 of the 91
    // stages) go through shared memory and a CTA barrier.  (The shared-memory-only network below paid a barrier and a dependent
    // LDS -> compare -> STS chain per stage: 84 us per 32 frames on 32 SMs; ncu source view of round 2.)
    for (int i = sort_n + t; i < TOPK_CAP; i += TOPK_NT) skeys[i] = 0;
    __syncthreads();
    u64 k[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(skeys + 8 * t + 2 * c);
      k[2 * c] = v.x; k[2 * c + 1] = v.y;
    }
    __syncthreads();
    const int lane = t & 31;
    // size 2, size 4: the direction depends on the element's own index bits
    tk_cx(k[0], k[1], true); tk_cx(k[2], k[3], false); tk_cx(k[4], k[5], true); tk_cx(k[6], k[7], false);
    tk_cx(k[0], k[2], true); tk_cx(k[1], k[3], true); tk_cx(k[4], k[6], false); tk_cx(k[5], k[7], false);
    tk_cx(k[0], k[1], true); tk_cx(k[2], k[3], true); tk_cx(k[4], k[5], false); tk_cx(k[6], k[7], false);
#pragma unroll 1
    for (int size = 8; size <= TOPK_CAP; size <<= 1) {
      const bool desc = ((8 * t) & size) == 0;
#pragma unroll 1
      for (int stride = size >> 1; stride >= 8; stride >>= 1) {
        const int tm = stride >> 3;                    // partner thread = t ^ tm
        const bool keepmax = ((t & tm) == 0) == desc;  // the lower element of a pair keeps the maximum in a descending run
        u64 o[8];
        if (tm >= 32) {
#pragma unroll
          for (int c = 0; c < 4; ++c) *tk_slot(skeys, t, c) = make_ulonglong2(k[2 * c], k[2 * c + 1]);
          __syncthreads();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const ulonglong2 v = *tk_slot(skeys, t ^ tm, c);
            o[2 * c] = v.x; o[2 * c + 1] = v.y;
          }
          __syncthreads();
        } else {
#pragma unroll
          for (int r = 0; r < 8; ++r) o[r] = __shfl_xor_sync(0xffffffffu, k[r], tm);
        }
        tk_keep(k, o, keepmax);
      }
      tk_local(k, desc);
    }
    (void)lane;
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<ulonglong2*>(skeys + 8 * t + 2 * c) = make_ulonglong2(k[2 * c], k[2 * c + 1]);
    __syncthreads();
  } else
  // bitonic sort, descending (small candidate sets: one pair per thread and stage)
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
// sum over the 16 lanes that share a keypoint (xor offsets 8 .. 1 stay inside a half warp)
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// channel-normalised corner: feats[y][x][:] / max(||.||_2, 1e-12)  (F.normalize(M1, dim=1), :273); this lane's 4 channels.
// The quotient is x * (1 / den) with an IEEE-rounded reciprocal: within 1.5 ulp of the reference's division, 1e-7 against a 1e-4
// tolerance, and one reciprocal per corner instead of four divisions per lane (the kernel is instruction bound).
__device__ __forceinline__ float4 corner_unit(const float* fmap, int mh, int mw, int x, int y, int sub) {
  // (no early return: the two keypoints of a warp may differ here and the shuffles below need all 32 lanes; a zero vector stays zero)
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (x >= 0 && x < mw && y >= 0 && y < mh) v = *reinterpret_cast<const float4*>(fmap + ((size_t)y * mw + x) * 64 + sub * 4);
  const float nrm = sqrtf(half_warp_sum(fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)))));
  const float inv = __frcp_rn(fmaxf(nrm, 1e-12f));
  return make_float4(__fmul_rn(v.x, inv), __fmul_rn(v.y, inv), __fmul_rn(v.z, inv), __fmul_rn(v.w, inv));
}

// 16 lanes per keypoint (4 channels each: a corner is one 256-byte line read as 16 x LDG.128), two keypoints per warp
__global__ void __launch_bounds__(256) describe_kernel(const float* feats, int H, int W, int topk, const int32_t* n_valid,
                                                       const float* kpt_xy, float* desc) {
  const int b = blockIdx.y;
  const int sub = threadIdx.x & 15;
  const int i = blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4);
  const bool live = i < topk && i < n_valid[b];       // (the branch is uniform over the 16 lanes of a keypoint; shuffles stay inside them)
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  // every lane of the warp runs the arithmetic (the half-warp shuffles need all 32 lanes converged); dead keypoints read nothing
  const int ii = live ? i : 0;
  const int mh = H >> 3, mw = W >> 3;
  const float* fmap = feats + (size_t)b * mh * mw * 64;
  const float2 kp = live ? *reinterpret_cast<const float2*>(kpt_xy + ((size_t)b * topk + ii) * 2) : make_float2(0.f, 0.f);
  const int px = (int)kp.x, py = (int)kp.y;
  const float sx = grid_src(px, W, mw), sy = grid_src(py, H, mh);
  const float x0f = floorf(sx), y0f = floorf(sy);
  const int x0 = live ? (int)x0f : -4, y0 = (int)y0f;  // dead: every corner is out of the map
  const float w = __fsub_rn(sx, x0f), e = __fsub_rn(1.0f, w), n = __fsub_rn(sy, y0f), s = __fsub_rn(1.0f, n);
  const float4 nw = corner_unit(fmap, mh, mw, x0, y0, sub);
  const float4 ne = corner_unit(fmap, mh, mw, x0 + 1, y0, sub);
  const float4 sw = corner_unit(fmap, mh, mw, x0, y0 + 1, sub);
  const float4 se = corner_unit(fmap, mh, mw, x0 + 1, y0 + 1, sub);
  const float wnw = __fmul_rn(s, e), wne = __fmul_rn(s, w), wsw = __fmul_rn(n, e), wse = __fmul_rn(n, w);
  float4 v;
  v.x = bilerp4(nw.x, ne.x, sw.x, se.x, wnw, wne, wsw, wse);
  v.y = bilerp4(nw.y, ne.y, sw.y, se.y, wnw, wne, wsw, wse);
  v.z = bilerp4(nw.z, ne.z, sw.z, se.z, wnw, wne, wsw, wse);
  v.w = bilerp4(nw.w, ne.w, sw.w, se.w, wnw, wne, wsw, wse);
  const float nrm = sqrtf(half_warp_sum(fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)))));
  const float inv = __frcp_rn(fmaxf(nrm, 1e-12f));
  if (live) o = make_float4(__fmul_rn(v.x, inv), __fmul_rn(v.y, inv), __fmul_rn(v.z, inv), __fmul_rn(v.w, inv));
  if (i < topk) *reinterpret_cast<float4*>(desc + ((size_t)b * topk + i) * 64 + sub * 4) = o;
}

cudaError_t launch_post(Ctx* c, int topk, float nms_thr, int32_t* d_nvalid, float* d_xy, float* d_score, float* d_desc) {
  const int H = c->H, W = c->W;
  dim3 g1((W + NMS_TW - 1) / NMS_TW, (H + NMS_TH - 1) / NMS_TH, c->B);
  prof_begin(c, P_NMS);
  nms_score_kernel<<<g1, NMS_NT, 0, c->stream>>>(c->k1h, c->act[L_HM_2], H, W, nms_thr, c->cand, c->cand_count);
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
  dim3 g3((topk + 15) / 16, c->B);
  prof_begin(c, P_DESCRIBE);
  describe_kernel<<<g3, 256, 0, c->stream>>>(c->act[L_F_2], H, W, topk, d_nvalid, d_xy, d_desc);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

}  // namespace xfb
