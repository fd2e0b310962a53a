// tmem_bench.cu -- microbenchmarks of the sm_100a tensor-memory path that the matcher / conv kernels are shaped by:
//   (1) tcgen05.ld 32x32b latency (one warp, dependent ld -> wait) and throughput (W warps, several lds in flight)
//   (2) tcgen05.mma kind::f16 issue-to-completion rate, M = 128, N in {128, 256}, A from shared memory (SS) or tensor memory (TS)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_bench tools/tmem_bench.cu ; run on one B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "../xfeatslam_b200/csrc/tc_ptx.cuh"
using namespace xfb;

__device__ __forceinline__ void ld_x32(uint32_t taddr, uint32_t* u) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
      "%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
        "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]),
        "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]),
        "=r"(u[31])
      : "r"(taddr));
}
__device__ __forceinline__ void ld_x16(uint32_t taddr, uint32_t* u) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                 "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: ld x32 + wait, dependent (latency).  mode 1: 2 x (ld x32) in flight per wait.  mode 2: 4 x (ld x16) per wait.
template <int MODE>
__global__ void __launch_bounds__(512) ld_kernel(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&s_tmem, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tq = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128) % 512u;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      uint32_t u[32];
      ld_x32(tq + (i & 3) * 32, u);
      ld_wait();
      acc += u[0] ^ u[31];
    } else if (MODE == 1) {
      uint32_t u[32], v[32];
      ld_x32(tq + (i & 1) * 64, u);
      ld_x32(tq + (i & 1) * 64 + 32, v);
      ld_wait();
      acc += u[0] ^ v[31];
    } else {
      uint32_t u[16], v[16], w[16], x[16];
      ld_x16(tq, u); ld_x16(tq + 16, v); ld_x16(tq + 32, w); ld_x16(tq + 48, x);
      ld_wait();
      acc += u[0] ^ v[15] ^ w[3] ^ x[7];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(s_tmem, 512u);
}

__host__ __device__ constexpr uint32_t idesc_f16(uint32_t N) { return (1u << 4) | ((N >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd),
               "r"(idesc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t at, uint64_t bd, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(at), "l"(bd),
               "r"(idesc)
               : "memory");
}

// One thread issues `iters` MMAs (K = 16 each) back to back, commits, waits.  SWZ: 0 = no-swizzle K-major canonical layout
// (LBO 2048 / 4096, SBO 128), operands are whatever bytes sit in shared memory (zeros).
template <int N, bool TS>
__global__ void __launch_bounds__(128) mma_kernel(int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  for (int i = threadIdx.x; i < (4096 + N * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc(&s_tmem, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint64_t ad = umma_desc_kmajor(smem_u32(smem), 2048, 128);                 // A: 128 rows x 16 fp16
    const uint64_t bd = umma_desc_kmajor(smem_u32(smem + 4096), N * 16, 128);        // B: N rows x 16 fp16
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (TS) mma_ts(s_tmem + (i & 1) * 256, s_tmem + 300, bd, idesc_f16(N));
      else mma_ss(s_tmem + (i & 1) * 256, ad, bd, idesc_f16(N));
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(s_tmem, 512u); }
}

int main() {
  long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 64); cudaMalloc(&d_sink, 64);
  long long h = 0;
  const int iters = 4096;
  auto report_ld = [&](const char* name, int warps, double bytes_per_iter_per_warp) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-34s warps %2d : %8.1f clk/iter  -> %7.1f B/clk/SM  (%s)\n", name, warps, (double)h / iters, bytes_per_iter_per_warp * warps * iters / (double)h,
           cudaGetErrorString(e));
  };
  for (int warps : {1, 4, 8, 16}) {
    ld_kernel<0><<<148, warps * 32>>>(iters, d_out, d_sink); report_ld("ld x32 + wait (dependent)", warps, 4096.0);
    ld_kernel<1><<<148, warps * 32>>>(iters, d_out, d_sink); report_ld("2 x ld x32 per wait", warps, 8192.0);
    ld_kernel<2><<<148, warps * 32>>>(iters, d_out, d_sink); report_ld("4 x ld x16 per wait", warps, 8192.0);
  }
  auto report_mma = [&](const char* name, int N) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    const double clk = (double)h / iters;
    printf("%-34s : %7.1f clk/MMA -> %6.0f MAC/clk/SM (peak 4096)  (%s)\n", name, clk, 128.0 * N * 16 / clk, cudaGetErrorString(e));
  };
  cudaFuncSetAttribute(mma_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(mma_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(mma_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(mma_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  mma_kernel<128, false><<<148, 128, 65536>>>(iters, d_out); report_mma("mma f16 M128 N128 K16 SS", 128);
  mma_kernel<256, false><<<148, 128, 65536>>>(iters, d_out); report_mma("mma f16 M128 N256 K16 SS", 256);
  mma_kernel<128, true><<<148, 128, 65536>>>(iters, d_out); report_mma("mma f16 M128 N128 K16 TS", 128);
  mma_kernel<256, true><<<148, 128, 65536>>>(iters, d_out); report_mma("mma f16 M128 N256 K16 TS", 256);
  return 0;
}
