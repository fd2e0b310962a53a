// match.cu -- 64-D descriptor distances and brute-force nearest / second-nearest search.
//
// Distance = ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2242-2250, XFeat branch):
//     int(float(cv::norm(a, b, NORM_L2SQR)) * 512)
// with cv::norm restated as: fp32 subtract, fp64 multiply-accumulate in index order 0..63 (see
// oracle/matcher_oracle.c).  (double)d*(double)d is exact, so fma == mul+add here and the result is
// bit-identical to the oracle.  Best / second-best follow the reference's scan rule
// (src/ORBmatcher.cc:476-486): strict '<', candidates visited in ascending index order, i.e. the
// total order (distance, index) -- which makes the reduction order-independent on the GPU.
//
// v1: exact FP64 SIMT tiles (64 x 64 pairs per CTA, 4 x 4 per thread).
#include <cuda_runtime.h>

#include "xfb_internal.h"

namespace xfb {

constexpr int MT = 64;       // tile edge (pairs)
constexpr int MLD = 65;      // padded row length in shared memory

struct Top2 { int d1, idx, d2; };

__device__ __forceinline__ void top2_push(Top2& s, int d, int j) {
  if (d < s.d1) { s.d2 = s.d1; s.d1 = d; s.idx = j; }
  else if (d < s.d2) { s.d2 = d; }
}
__device__ __forceinline__ Top2 top2_merge(const Top2& a, const Top2& b) {
  Top2 r;
  if (a.d1 < b.d1 || (a.d1 == b.d1 && a.idx < b.idx)) { r.d1 = a.d1; r.idx = a.idx; r.d2 = min(a.d2, b.d1); }
  else { r.d1 = b.d1; r.idx = b.idx; r.d2 = min(b.d2, a.d1); }
  return r;
}

__device__ __forceinline__ void load_tile(float* s, const float* g, int row0, int n, int t) {
  // 64 rows x 64 floats; rows >= n are zero-filled
  for (int i = t; i < MT * 16; i += 256) {
    const int r = i >> 4, q = i & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < n) v = *reinterpret_cast<const float4*>(g + (size_t)(row0 + r) * 64 + q * 4);
    s[r * MLD + q * 4 + 0] = v.x; s[r * MLD + q * 4 + 1] = v.y; s[r * MLD + q * 4 + 2] = v.z; s[r * MLD + q * 4 + 3] = v.w;
  }
}

template <bool MATRIX>
__global__ void __launch_bounds__(256) dist_tile_kernel(const float* A, int n1, const float* B, int n2, const int32_t* n1p,
                                                        const int32_t* n2p, const int32_t* ga, const int32_t* gb, int init,
                                                        int32_t* out_matrix, int32_t* rowpart, int32_t* colpart) {
  if (n1p) n1 = min(n1, *n1p);   // device-resident counts (xfb_match_frames): n1/n2 are the capacities
  if (n2p) n2 = min(n2, *n2p);
  if ((int)blockIdx.y * MT >= n1 || (int)blockIdx.x * MT >= n2) return;
  __shared__ float sA[MT * MLD];
  __shared__ float sB[MT * MLD];
  __shared__ unsigned long long sCol[16][MT];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int row0 = blockIdx.y * MT, col0 = blockIdx.x * MT;
  load_tile(sA, A, row0, n1, t);
  load_tile(sB, B, col0, n2, t);
  __syncthreads();
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 4
  for (int k = 0; k < 64; ++k) {
    float a[4], bb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = sA[(ty * 4 + i) * MLD + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) bb[j] = sB[(tx * 4 + j) * MLD + k];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = a[i] - bb[j];
        const double dd = (double)d;
        acc[i][j] = fma(dd, dd, acc[i][j]);
      }
  }
  int dist[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dist[i][j] = (int)(__double2float_rn(acc[i][j]) * 512.0f);

  if (MATRIX) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = row0 + ty * 4 + i;
      if (r >= n1) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cidx = col0 + tx * 4 + j;
        if (cidx < n2) out_matrix[(size_t)r * n2 + cidx] = dist[i][j];
      }
    }
    return;
  }

  int gai[4], gbj[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) gai[i] = (ga && row0 + ty * 4 + i < n1) ? ga[row0 + ty * 4 + i] : 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) gbj[j] = (gb && col0 + tx * 4 + j < n2) ? gb[col0 + tx * 4 + j] : 0;
  const bool grouped = (ga != nullptr) && (gb != nullptr);

  // row-wise: this thread's 4 columns, then the 16 threads sharing a row (tx = lane & 15)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Top2 s; s.d1 = init; s.idx = -1; s.d2 = init;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cidx = col0 + tx * 4 + j;
      if (cidx < n2 && (!grouped || gai[i] == gbj[j])) top2_push(s, dist[i][j], cidx);
    }
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
      Top2 o;
      o.d1 = __shfl_xor_sync(0xffffffffu, s.d1, off);
      o.idx = __shfl_xor_sync(0xffffffffu, s.idx, off);
      o.d2 = __shfl_xor_sync(0xffffffffu, s.d2, off);
      s = top2_merge(s, o);
    }
    const int r = row0 + ty * 4 + i;
    if (tx == 0 && r < n1) {
      int32_t* dst = rowpart + ((size_t)r * gridDim.x + blockIdx.x) * 3;
      dst[0] = s.d1; dst[1] = s.idx; dst[2] = s.d2;
    }
  }
  // column-wise best (distance, row index): this thread's 4 rows, then the 16 ty groups via smem
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int bd = init, bi = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = row0 + ty * 4 + i;
      if (r < n1 && (!grouped || gai[i] == gbj[j]) && dist[i][j] < bd) { bd = dist[i][j]; bi = r; }
    }
    sCol[ty][tx * 4 + j] = ((unsigned long long)(unsigned int)bd << 32) | (unsigned int)bi;
  }
  __syncthreads();
  if (t < MT) {
    // ty ascending == row index ascending, strict '<' keeps the lowest row on ties
    int bd = init, bi = -1;
    for (int g = 0; g < 16; ++g) {
      const unsigned long long p = sCol[g][t];
      const int d = (int)(unsigned int)(p >> 32), i = (int)(unsigned int)(p & 0xffffffffu);
      if (d < bd) { bd = d; bi = i; }
    }
    const int cidx = col0 + t;
    if (cidx < n2) {
      int32_t* dst = colpart + ((size_t)cidx * gridDim.y + blockIdx.y) * 2;
      dst[0] = bd; dst[1] = bi;
    }
  }
}

__global__ void match_merge_kernel(const int32_t* rowpart, const int32_t* colpart, int n1cap, int n2cap, const int32_t* n1p,
                                   const int32_t* n2p, int nct_stride, int nrt_stride, int init, int32_t* bi, int32_t* bd,
                                   int32_t* sd, int32_t* ri, int32_t* rd) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int n1 = n1p ? min(n1cap, *n1p) : n1cap, n2 = n2p ? min(n2cap, *n2p) : n2cap;
  // tiles that were actually computed (a tile kernel CTA exits early beyond n1 / n2)
  const int nct_live = (n1 > 0 && n2 > 0) ? (n2 + MT - 1) / MT : 0, nrt_live = (n1 > 0 && n2 > 0) ? (n1 + MT - 1) / MT : 0;
  const int nct = nct_stride, nrt = nrt_stride;
  if (g < n1cap) {
    Top2 s; s.d1 = init; s.idx = -1; s.d2 = init;
    for (int c = 0; c < (g < n1 ? nct_live : 0); ++c) {
      const int32_t* p = rowpart + ((size_t)g * nct + c) * 3;
      Top2 o; o.d1 = p[0]; o.idx = p[1]; o.d2 = p[2];
      if (o.idx >= 0 || o.d2 < init) s = top2_merge(s, o);
    }
    if (bi) bi[g] = s.idx;
    if (bd) bd[g] = s.d1;
    if (sd) sd[g] = s.d2;
  }
  if (g < n2cap) {
    int bdist = init, bidx = -1;
    for (int r = 0; r < (g < n2 ? nrt_live : 0); ++r) {
      const int32_t* p = colpart + ((size_t)g * nrt + r) * 2;
      if (p[0] < bdist) { bdist = p[0]; bidx = p[1]; }
    }
    if (ri) ri[g] = bidx;
    if (rd) rd[g] = bdist;
  }
}

cudaError_t launch_distance_matrix(Ctx* c, const float* dA, int n1, const float* dB, int n2, int32_t* d_out) {
  if (n1 <= 0 || n2 <= 0) return cudaSuccess;
  dim3 grid((n2 + MT - 1) / MT, (n1 + MT - 1) / MT);
  prof_begin(c, P_DIST_MATRIX);
  dist_tile_kernel<true><<<grid, 256, 0, c->stream>>>(dA, n1, dB, n2, nullptr, nullptr, nullptr, nullptr, 0, d_out, nullptr, nullptr);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

cudaError_t launch_match(Ctx* c, const float* dA, int n1, const float* dB, int n2, const int32_t* ga, const int32_t* gb, int init,
                         int32_t* bi, int32_t* bd, int32_t* sd, int32_t* ri, int32_t* rd, const int32_t* n1p, const int32_t* n2p) {
  if (n1 <= 0 && n2 <= 0) return cudaSuccess;
  const int nct = (n2 + MT - 1) / MT, nrt = (n1 + MT - 1) / MT;
  if (n1 > 0 && n2 > 0) {
    dim3 grid(nct, nrt);
    prof_begin(c, P_MATCH_TILE);
    dist_tile_kernel<false><<<grid, 256, 0, c->stream>>>(dA, n1, dB, n2, n1p, n2p, ga, gb, init, nullptr, c->m_rowpart, c->m_colpart);
    prof_end(c);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const int n = n1 > n2 ? n1 : n2;
  prof_begin(c, P_MATCH_MERGE);
  match_merge_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->m_rowpart, c->m_colpart, n1, n2, n1p, n2p, nct, nrt, init, bi, bd, sd, ri,
                                                             rd);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

}  // namespace xfb
