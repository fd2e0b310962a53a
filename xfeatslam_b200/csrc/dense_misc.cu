// dense_misc.cu -- the non-convolution dense stages of the XFeat forward:
//   prep      : u8 -> f32/255 (XFextractor::parseInput, src/XFextractor.cc:161-176), bilinear resize to
//               a multiple of 32 (preprocessTensor, :182-202), InstanceNorm2d(1) (src/XFeat.cc:148-149),
//               AvgPool2d(4,4) of skip1 (src/XFeat.cc:36-39)
//   pyramid   : x3 + up2(x4) + up4(x5), bilinear align_corners=false (src/XFeat.cc:159-166)
//   heatmap   : heatmap_head.2 (1x1 64->1 + bias) + Sigmoid (src/XFeat.cc:78-83)
//   keypoints : keypoint_head.3 (1x1 64->65 + bias) + softmax over 65 + drop dustbin + 8x8 fold
//               (src/XFeat.cc:85-90, XFextractor::getKptsHeatmap src/XFextractor.cc:204-217)
#include <cuda_runtime.h>
#include <stdint.h>

#include "xfb_internal.h"

namespace xfb {

// ------------------------------------------------------------------------------------------------
// prep pass 1: x_pre and per-frame sum / sum of squares (fixed-order FP64 fold by the last CTA)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pre_value(const uint8_t* img, int stride, int in_h, int in_w, int H, int W, int y, int x,
                                           float sh, float sw) {
  if (in_h == H && in_w == W) return (float)img[(size_t)y * stride + x] / 255.0f;
  // at::upsample_bilinear2d, align_corners=false: src = scale*(dst+0.5)-0.5 clamped at 0
  float sy = sh * ((float)y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
  float sx = sw * ((float)x + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
  int y0 = (int)floorf(sy); if (y0 > in_h - 1) y0 = in_h - 1;
  int x0 = (int)floorf(sx); if (x0 > in_w - 1) x0 = in_w - 1;
  float ly = fminf(fmaxf(sy - (float)y0, 0.f), 1.f);
  float lx = fminf(fmaxf(sx - (float)x0, 0.f), 1.f);
  const int y1 = y0 + (y0 < in_h - 1 ? 1 : 0), x1 = x0 + (x0 < in_w - 1 ? 1 : 0);
  const float v00 = (float)img[(size_t)y0 * stride + x0] / 255.0f, v01 = (float)img[(size_t)y0 * stride + x1] / 255.0f;
  const float v10 = (float)img[(size_t)y1 * stride + x0] / 255.0f, v11 = (float)img[(size_t)y1 * stride + x1] / 255.0f;
  const float wy0 = 1.f - ly, wx0 = 1.f - lx;
  return wy0 * (wx0 * v00 + lx * v01) + ly * (wx0 * v10 + lx * v11);
}

constexpr int PREP_NT = 256;
constexpr int PREP_PIX = 2048;  // pixels per CTA

__global__ void __launch_bounds__(PREP_NT) prep_stats_kernel(const uint8_t* gray, size_t frame_stride, int stride, int in_h,
                                                             int in_w, int H, int W, float* xraw, double* part,
                                                             unsigned int* ticket, float* mean_out, float* rstd_out) {
  const int b = blockIdx.y, t = threadIdx.x;
  const uint8_t* img = gray + (size_t)b * frame_stride;
  const float sh = (float)in_h / (float)H, sw = (float)in_w / (float)W;
  const int npix = H * W;
  const int base = blockIdx.x * PREP_PIX;
  float s1 = 0.f, s2 = 0.f;
  // the 256 possible values of (float)u8 / 255.0f, each the correctly rounded quotient the reference computes: a table lookup instead
  // of four IEEE divisions per 4 pixels (the kernel was instruction bound on them)
  __shared__ float s_lut[256];
  s_lut[t] = (float)t / 255.0f;
  __syncthreads();
  static_assert(PREP_NT == 256, "one table entry per thread");
  if (in_h == H && in_w == W && (stride & 3) == 0 && (frame_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(gray) & 3) == 0) {
    // no resize (the usual case): 4 pixels per thread and step -- one 32-bit load, one 16-byte store (W is a multiple of 32,
    // so a group of 4 never crosses a row)
    for (int i = t * 4; i < PREP_PIX; i += PREP_NT * 4) {
      const int p = base + i;
      if (p < npix) {
        const int y = p / W, x = p - y * W;
        const uchar4 u = *reinterpret_cast<const uchar4*>(img + (size_t)y * stride + x);
        const float4 v = make_float4(s_lut[u.x], s_lut[u.y], s_lut[u.z], s_lut[u.w]);
        *reinterpret_cast<float4*>(xraw + (size_t)b * npix + p) = v;
        s1 += v.x; s2 = fmaf(v.x, v.x, s2);
        s1 += v.y; s2 = fmaf(v.y, v.y, s2);
        s1 += v.z; s2 = fmaf(v.z, v.z, s2);
        s1 += v.w; s2 = fmaf(v.w, v.w, s2);
      }
    }
  } else {
    for (int i = t; i < PREP_PIX; i += PREP_NT) {
      const int p = base + i;
      if (p < npix) {
        const int y = p / W, x = p - y * W;
        const float v = pre_value(img, stride, in_h, in_w, H, W, y, x, sh, sw);
        xraw[(size_t)b * npix + p] = v;
        s1 += v;
        s2 = fmaf(v, v, s2);
      }
    }
  }
  __shared__ double sred[PREP_NT / 32][2];
  __shared__ unsigned int s_last;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    s2 += __shfl_xor_sync(0xffffffffu, s2, off);
  }
  if ((t & 31) == 0) { sred[t >> 5][0] = (double)s1; sred[t >> 5][1] = (double)s2; }
  __syncthreads();
  const int tiles = gridDim.x;
  double* part_b = part + (size_t)b * tiles * 2;
  if (t == 0) {
    double d1 = 0.0, d2 = 0.0;
    for (int w = 0; w < PREP_NT / 32; ++w) { d1 += sred[w][0]; d2 += sred[w][1]; }
    part_b[blockIdx.x * 2] = d1;
    part_b[blockIdx.x * 2 + 1] = d2;
    __threadfence();
    const unsigned int prev = atomicAdd(ticket + b * XFB_TICKET_STRIDE, 1u);
    s_last = (prev == (unsigned int)(tiles - 1)) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // fixed-order fold: thread t sums tiles t, t+NT, ...; then thread 0 sums the NT slices in order
  __shared__ double sfold[PREP_NT][2];
  double d1 = 0.0, d2 = 0.0;
  for (int i = t; i < tiles; i += PREP_NT) { d1 += __ldcg(part_b + 2 * i); d2 += __ldcg(part_b + 2 * i + 1); }
  sfold[t][0] = d1; sfold[t][1] = d2;
  __syncthreads();
  if (t == 0) {
    d1 = 0.0; d2 = 0.0;
    for (int i = 0; i < PREP_NT; ++i) { d1 += sfold[i][0]; d2 += sfold[i][1]; }
    const double n = (double)npix;
    const double mean = d1 / n;
    double var = d2 / n - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_out[b] = (float)mean;
    rstd_out[b] = (float)(1.0 / sqrt(var + 1e-5));
    ticket[b * XFB_TICKET_STRIDE] = 0u;
  }
}

// prep pass 2: xn = (x - mean) * rstd and the 4x4 average pool of xn.  One thread per 4x4 cell.
__global__ void __launch_bounds__(256) prep_norm_kernel(const float* xraw, const float* mean, const float* rstd, int H, int W,
                                                        float* xn, float* avg4) {
  const int b = blockIdx.z;
  const int cx = blockIdx.x * blockDim.x + threadIdx.x;
  const int cy = blockIdx.y * blockDim.y + threadIdx.y;
  const int W4 = W >> 2, H4 = H >> 2;
  if (cx >= W4 || cy >= H4) return;
  const float m = mean[b], r = rstd[b];
  const float* src = xraw + (size_t)b * H * W;
  float* dst = xn + (size_t)b * H * W;
  float s = 0.f;
#pragma unroll
  for (int dy = 0; dy < 4; ++dy) {
    const size_t o = (size_t)(cy * 4 + dy) * W + cx * 4;
    float4 v = *reinterpret_cast<const float4*>(src + o);
    v.x = (v.x - m) * r; v.y = (v.y - m) * r; v.z = (v.z - m) * r; v.w = (v.w - m) * r;
    *reinterpret_cast<float4*>(dst + o) = v;
    s += v.x; s += v.y; s += v.z; s += v.w;
  }
  avg4[((size_t)b * H4 + cy) * W4 + cx] = s / 16.0f;
}

cudaError_t launch_prep(Ctx* c, const uint8_t* d_gray, size_t frame_stride, int stride) {
  const int npix = c->H * c->W;
  dim3 g1((npix + PREP_PIX - 1) / PREP_PIX, c->B);
  prof_begin(c, P_PREP_STATS);
  prep_stats_kernel<<<g1, PREP_NT, 0, c->stream>>>(d_gray, frame_stride, stride, c->in_h, c->in_w, c->H, c->W, c->xraw, c->part,
                                                   c->ticket, c->in_mean, c->in_rstd);
  prof_end(c);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 blk(32, 8);
  dim3 g2(((c->W >> 2) + 31) / 32, ((c->H >> 2) + 7) / 8, c->B);
  prof_begin(c, P_PREP_NORM);
  prep_norm_kernel<<<g2, blk, 0, c->stream>>>(c->xraw, c->in_mean, c->in_rstd, c->H, c->W, c->xn, c->avg4);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// pyramid fusion input: relu(bn(x3)) + up2(relu(bn(x4))) + up4(relu(bn(x5)))
// ------------------------------------------------------------------------------------------------
struct UpTap { int i0, i1; float w0, w1; };
__device__ __forceinline__ UpTap up_tap(int dst, int in_size, float scale) {
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  int i0 = (int)floorf(s);
  if (i0 > in_size - 1) i0 = in_size - 1;
  float l = fminf(fmaxf(s - (float)i0, 0.f), 1.f);
  UpTap t;
  t.i0 = i0; t.i1 = i0 + (i0 < in_size - 1 ? 1 : 0); t.w0 = 1.f - l; t.w1 = l;
  return t;
}

__device__ __forceinline__ float4 bnrelu4(const float* p, const float* mean, const float* rstd) {
  float4 v = *reinterpret_cast<const float4*>(p);
  const float4 m = *reinterpret_cast<const float4*>(mean);
  const float4 r = *reinterpret_cast<const float4*>(rstd);
  v.x = fmaxf((v.x - m.x) * r.x, 0.f); v.y = fmaxf((v.y - m.y) * r.y, 0.f);
  v.z = fmaxf((v.z - m.z) * r.z, 0.f); v.w = fmaxf((v.w - m.w) * r.w, 0.f);
  return v;
}

__device__ __forceinline__ float4 up_sample(const float* src, int hs, int ws, int y, int x, int c, float sh, float sw,
                                            const float* mean, const float* rstd) {
  const UpTap ty = up_tap(y, hs, sh), tx = up_tap(x, ws, sw);
  const float4 v00 = bnrelu4(src + ((size_t)ty.i0 * ws + tx.i0) * 64 + c, mean, rstd);
  const float4 v01 = bnrelu4(src + ((size_t)ty.i0 * ws + tx.i1) * 64 + c, mean, rstd);
  const float4 v10 = bnrelu4(src + ((size_t)ty.i1 * ws + tx.i0) * 64 + c, mean, rstd);
  const float4 v11 = bnrelu4(src + ((size_t)ty.i1 * ws + tx.i1) * 64 + c, mean, rstd);
  float4 o;
  o.x = ty.w0 * (tx.w0 * v00.x + tx.w1 * v01.x) + ty.w1 * (tx.w0 * v10.x + tx.w1 * v11.x);
  o.y = ty.w0 * (tx.w0 * v00.y + tx.w1 * v01.y) + ty.w1 * (tx.w0 * v10.y + tx.w1 * v11.y);
  o.z = ty.w0 * (tx.w0 * v00.z + tx.w1 * v01.z) + ty.w1 * (tx.w0 * v10.z + tx.w1 * v11.z);
  o.w = ty.w0 * (tx.w0 * v00.w + tx.w1 * v01.w) + ty.w1 * (tx.w0 * v10.w + tx.w1 * v11.w);
  return o;
}

__global__ void __launch_bounds__(256) pyramid_kernel(const float* x3, const float* x4, const float* x5, const float* m3,
                                                      const float* r3, const float* m4, const float* r4, const float* m5,
                                                      const float* r5, int h3, int w3, int h4, int w4, int h5, int w5,
                                                      float* out) {
  const int b = blockIdx.y;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;  // (pixel, channel quad)
  const int npix = h3 * w3;
  if (gid >= npix * 16) return;
  const int c = (gid & 15) * 4;
  const int p = gid >> 4;
  const int y = p / w3, x = p - y * w3;
  const float4 a = bnrelu4(x3 + ((size_t)b * npix + p) * 64 + c, m3 + b * 64 + c, r3 + b * 64 + c);
  const float4 u4 = up_sample(x4 + (size_t)b * h4 * w4 * 64, h4, w4, y, x, c, (float)h4 / (float)h3, (float)w4 / (float)w3,
                              m4 + b * 64 + c, r4 + b * 64 + c);
  const float4 u5 = up_sample(x5 + (size_t)b * h5 * w5 * 64, h5, w5, y, x, c, (float)h5 / (float)h3, (float)w5 / (float)w3,
                              m5 + b * 64 + c, r5 + b * 64 + c);
  float4 o;
  o.x = (a.x + u4.x) + u5.x; o.y = (a.y + u4.y) + u5.y; o.z = (a.z + u4.z) + u5.z; o.w = (a.w + u4.w) + u5.w;
  *reinterpret_cast<float4*>(out + ((size_t)b * npix + p) * 64 + c) = o;
}

cudaError_t launch_pyramid(Ctx* c) {
  const int h3 = c->H >> 3, w3 = c->W >> 3, h4 = c->H >> 4, w4 = c->W >> 4, h5 = c->H >> 5, w5 = c->W >> 5;
  dim3 grid((h3 * w3 * 16 + 255) / 256, c->B);
  prof_begin(c, P_PYRAMID);
  pyramid_kernel<<<grid, 256, 0, c->stream>>>(c->act[L_B3_2], c->act[L_B4_2], c->act[L_B5_3], c->bn[L_B3_2].mean, c->bn[L_B3_2].rstd,
                                              c->bn[L_B4_2].mean, c->bn[L_B4_2].rstd, c->bn[L_B5_3].mean, c->bn[L_B5_3].rstd, h3, w3, h4,
                                              w4, h5, w5, c->pyr);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// heatmap_head.2 + sigmoid: 16 lanes per pixel (4 channels each: a pixel is one 256-byte line read as 16 x LDG.128), 4 pixels per
// thread with their loads in flight together, 4 shuffle steps per dot product.  (The first version used a whole warp per pixel.)
// ------------------------------------------------------------------------------------------------
constexpr int HO_PIX = 4;      // pixels per 16-lane group
__global__ void __launch_bounds__(256) heatmap_out_kernel(const float* in, const float* mean, const float* rstd, const float* w,
                                                          const float* bias, int npix, float* out) {
  const int b = blockIdx.y;
  const int sub = threadIdx.x & 15, grp = threadIdx.x >> 4;                    // 16 groups per CTA
  const int p0 = (blockIdx.x * 16 + grp) * HO_PIX;
  const float4 m = *reinterpret_cast<const float4*>(mean + b * 64 + sub * 4);
  const float4 r = *reinterpret_cast<const float4*>(rstd + b * 64 + sub * 4);
  const float4 ww = *reinterpret_cast<const float4*>(w + sub * 4);
  float4 v[HO_PIX];
#pragma unroll
  for (int i = 0; i < HO_PIX; ++i) {
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p0 + i < npix) v[i] = *reinterpret_cast<const float4*>(in + ((size_t)b * npix + p0 + i) * 64 + sub * 4);
  }
  const float bs = bias[0];
#pragma unroll
  for (int i = 0; i < HO_PIX; ++i) {
    float s = fmaxf((v[i].x - m.x) * r.x, 0.f) * ww.x;
    s = fmaf(fmaxf((v[i].y - m.y) * r.y, 0.f), ww.y, s);
    s = fmaf(fmaxf((v[i].z - m.z) * r.z, 0.f), ww.z, s);
    s = fmaf(fmaxf((v[i].w - m.w) * r.w, 0.f), ww.w, s);
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (sub == 0 && p0 + i < npix) out[(size_t)b * npix + p0 + i] = 1.0f / (1.0f + expf(-(s + bs)));
  }
}

cudaError_t launch_heatmap_out(Ctx* c) {
  const int npix = (c->H >> 3) * (c->W >> 3);
  dim3 grid((npix + 16 * HO_PIX - 1) / (16 * HO_PIX), c->B);
  prof_begin(c, P_HEATMAP_OUT);
  heatmap_out_kernel<<<grid, 256, 0, c->stream>>>(c->act[L_HM_1], c->bn[L_HM_1].mean, c->bn[L_HM_1].rstd, c->w[L_HM_2], c->bias[L_HM_2],
                                                  npix, c->act[L_HM_2]);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// keypoint_head.3 + softmax(65) + drop dustbin + fold to full resolution.
// One warp per 8x8 cell: lane l owns logits l and l+32 (lane 0 also the dustbin, channel 64).
// ------------------------------------------------------------------------------------------------
constexpr int KP_WARPS = 8;
__global__ void __launch_bounds__(KP_WARPS * 32) keypoint_out_kernel(const float* in, const float* mean, const float* rstd,
                                                                     const float* w /*[64][65]*/, const float* bias /*[65]*/,
                                                                     int h, int wd, float* k1h) {
  __shared__ float sW[64 * 65];
  __shared__ float sB[65];
  __shared__ float sX[KP_WARPS][64];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 64 * 65; i += blockDim.x) sW[i] = w[i];
  for (int i = threadIdx.x; i < 65; i += blockDim.x) sB[i] = bias[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ncell = h * wd;
  const int W = wd * 8;
  for (int cell = blockIdx.x * KP_WARPS + warp; cell < ncell; cell += gridDim.x * KP_WARPS) {
    const float2 v = *reinterpret_cast<const float2*>(in + ((size_t)b * ncell + cell) * 64 + lane * 2);
    const float2 m = *reinterpret_cast<const float2*>(mean + b * 64 + lane * 2);
    const float2 r = *reinterpret_cast<const float2*>(rstd + b * 64 + lane * 2);
    __syncwarp();
    sX[warp][lane * 2] = fmaxf((v.x - m.x) * r.x, 0.f);
    sX[warp][lane * 2 + 1] = fmaxf((v.y - m.y) * r.y, 0.f);
    __syncwarp();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 8
    for (int ci = 0; ci < 64; ++ci) {
      const float x = sX[warp][ci];
      a0 = fmaf(x, sW[ci * 65 + lane], a0);
      a1 = fmaf(x, sW[ci * 65 + lane + 32], a1);
      a2 = fmaf(x, sW[ci * 65 + 64], a2);  // dustbin (same value in every lane)
    }
    a0 += sB[lane]; a1 += sB[lane + 32]; a2 += sB[64];
    float mx = fmaxf(fmaxf(a0, a1), a2);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    const float e0 = expf(a0 - mx), e1 = expf(a1 - mx), e2 = expf(a2 - mx);
    float sum = e0 + e1;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    sum += e2;
    // channel c -> pixel (8*Y + c/8, 8*X + c%8)
    const int Y = cell / wd, X = cell - Y * wd;
    float* dst = k1h + (size_t)b * (h * 8) * W + (size_t)(Y * 8) * W + X * 8;
    dst[(size_t)(lane >> 3) * W + (lane & 7)] = e0 / sum;
    dst[(size_t)((lane >> 3) + 4) * W + (lane & 7)] = e1 / sum;
  }
}

cudaError_t launch_keypoint_out(Ctx* c) {
  const int h = c->H >> 3, wd = c->W >> 3;
  int blocks = (h * wd + KP_WARPS - 1) / KP_WARPS;
  if (blocks > 600) blocks = 600;
  dim3 grid(blocks, c->B);
  prof_begin(c, P_KEYPOINT_OUT);
  keypoint_out_kernel<<<grid, KP_WARPS * 32, 0, c->stream>>>(c->act[L_KP_2], c->bn[L_KP_2].mean, c->bn[L_KP_2].rstd, c->w[L_KP_3],
                                                            c->bias[L_KP_3], h, wd, c->k1h);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

}  // namespace xfb
