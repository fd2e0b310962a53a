// pairs.cu -- ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2242-2250) for an explicit LIST of (row of A, row of B)
// pairs: the primitive behind every matcher of the reference whose candidate set is not "all pairs" -- the vocabulary-node
// gated scans (SearchByBoW :408-610 / :950-1090, SearchForTriangulation :1092-1331), the projected-window searches
// (SearchByProjection :42 / :612 / :719 / :1861 / :2074, Fuse :1333 / :1525, SearchBySim3 :1642) and
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:329-403).  The host replays (xfeatslam_b200/host/XFBmatcher.cc)
// build the list in the reference's visiting order, get all distances in ONE launch and then run the reference's
// sequential accept / reject logic over them.
//
// Byte work, HBM / L2-bound: 512 B of descriptor rows per pair (rows repeat, so mostly L1 / L2 hits), one thread per pair,
// fp32 subtract -> fp64 accumulate in index order -> (float) -> * 512 -> truncate: bit-exact w.r.t. oracle/matcher_oracle.c.
#include <cuda_runtime.h>

#include "xfb_internal.h"

namespace xfb {

__global__ void __launch_bounds__(256) distance_pairs_kernel(const float* __restrict__ A, int n1, const float* __restrict__ B, int n2,
                                                             const int32_t* __restrict__ ia, const int32_t* __restrict__ ib, int n_pairs,
                                                             int32_t* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int i = ia[p], j = ib[p];
  if (i < 0 || i >= n1 || j < 0 || j >= n2) { out[p] = -1; return; }   // (the host entry point rejects these up front)
  const float4* a = reinterpret_cast<const float4*>(A + (size_t)i * 64);
  const float4* b = reinterpret_cast<const float4*>(B + (size_t)j * 64);
  double s = 0.0;
#pragma unroll 8
  for (int kq = 0; kq < 16; ++kq) {
    const float4 x = a[kq], y = b[kq];
    float d;
    d = x.x - y.x; s = fma((double)d, (double)d, s);
    d = x.y - y.y; s = fma((double)d, (double)d, s);
    d = x.z - y.z; s = fma((double)d, (double)d, s);
    d = x.w - y.w; s = fma((double)d, (double)d, s);
  }
  out[p] = (int)(__double2float_rn(s) * 512.0f);
}

cudaError_t launch_distance_pairs(Ctx* c, const float* dA, int n1, const float* dB, int n2, const int32_t* d_ia, const int32_t* d_ib, int n_pairs,
                                  int32_t* d_out) {
  if (n_pairs <= 0) return cudaSuccess;
  prof_begin(c, P_DIST_PAIRS);
  distance_pairs_kernel<<<(n_pairs + 255) / 256, 256, 0, c->stream>>>(dA, n1, dB, n2, d_ia, d_ib, n_pairs, d_out);
  prof_end(c);
  c->launches++;
  return cudaGetLastError();
}

}  // namespace xfb
