// Microbenchmark: the DMMA / barrier skeleton of one n = 200 tile GEMM of admm_shared_tile_kernel (config 3), nothing else: 16 warps per
// CTA, one CTA per SM, 25 row-blocks spread 1 / 2 per warp exactly as the kernel spreads them, 50 mma.sync.m8n8k4.f64 per row-block (a
// dependent chain per row-block, the kernel's issue order), a CTA barrier after every GEMM.  No operator loads, no shared-memory loads.
// Prints cycles per GEMM for: the kernel's split (1,2,1,2,...), an even two-per-warp split (32 row-blocks), and one warp per
// sub-partition owning everything (6-7 chains per warp).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tile_gemm_shape tile_gemm_shape.cu && ./tile_gemm_shape
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int MAXNR>
__global__ void __launch_bounds__(512, 1) skeleton(int mode, int gemms, int ksteps, double *sink, long long *cycles) {
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5;
  int nr;
  if (mode == 0 || mode == 3) nr = ((warp + 1) * 25) / 16 - (warp * 25) / 16;      // the kernel's split of 25 row-blocks over 16 warps
  else if (mode == 1) nr = 2;                                          // 32 row-blocks, two per warp
  else nr = warp < 4 ? (warp == 3 ? 7 : 6) : 0;                        // one warp per sub-partition owns the sub-partition's 6 / 7 chains
  double acc[MAXNR][2];
#pragma unroll
  for (int r = 0; r < MAXNR; ++r) acc[r][0] = acc[r][1] = 0.0;
  const double a = 1e-3 * threadIdx.x, b = 1e-3;
  __syncthreads();
  const long long t0 = clock64();
  for (int g = 0; g < gemms; ++g) {
    if (mode == 3) {          // the kernel's split with compile-time chain counts (as gemm_run<NR> has them), k loop unrolled by 3 k-pairs
      if (nr == 2) {
#pragma unroll 3
        for (int k = 0; k < ksteps; k += 2) { dmma(acc[0], a, b); dmma(acc[1], a, b); dmma(acc[0], b, a); dmma(acc[1], b, a); }
      } else {
#pragma unroll 3
        for (int k = 0; k < ksteps; k += 2) { dmma(acc[0], a, b); dmma(acc[0], b, a); }
      }
    } else
    for (int k = 0; k < ksteps; k += 2) {
#pragma unroll
      for (int r = 0; r < MAXNR; ++r) if (r < nr) dmma(acc[r], a, b);
#pragma unroll
      for (int r = 0; r < MAXNR; ++r) if (r < nr) dmma(acc[r], b, a);
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  double s = 0.0;
#pragma unroll
  for (int r = 0; r < MAXNR; ++r) s += acc[r][0] + acc[r][1];
  if (s == 123.456) sink[0] = s;
  if (smem[threadIdx.x] == 123.456) sink[1] = 1.0;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double *sink; long long *cyc; cudaMalloc(&sink, 16); cudaMalloc(&cyc, 8 * 256);
  const int smem = 200 * 1024, gemms = 2000, ksteps = 50;
  cudaFuncSetAttribute(skeleton<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(skeleton<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char *names[4] = {"kernel split, compile-time chain counts (no predicated DMMAs)", "kernel split 1/2 row-blocks per warp (25 row-blocks, busiest sub-partition 7)", "two row-blocks per warp (32 row-blocks, 8 per sub-partition)",
                          "one warp per sub-partition with 6/6/6/7 chains"};
  for (int grid : {1, sms})
    for (int mi = 0; mi < 4; ++mi) {
      const int mode = mi == 0 ? 3 : mi - 1;
      if (mode < 2) skeleton<2><<<grid, 512, smem>>>(mode, gemms, ksteps, sink, cyc); else skeleton<7><<<grid, 512, smem>>>(mode, gemms, ksteps, sink, cyc);
      cudaDeviceSynchronize();
      long long h[256]; cudaMemcpy(h, cyc, 8 * grid, cudaMemcpyDeviceToHost);
      double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
      const int busiest = mode == 1 ? 8 : 7;
      printf("{\"ctas\": %d, \"split\": \"%s\", \"cycles_per_gemm\": %.0f, \"dmma_floor_cycles\": %d, \"cycles_per_dmma_busiest_subpartition\": %.1f}\n",
             grid, names[mi], mean / gemms, busiest * ksteps * 16, mean / gemms / (busiest * ksteps));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
