// dmma_issue.cu -- how fast can ONE warp issue independent mma.sync.m8n8k4.f64, and how does the rate change with
// 1, 2, 4 warps per SM sub-partition?  (design input for the small-QP DMMA kernels: latency per ADMM iteration)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_issue dmma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void dmma_indep(double *out, int iters, long long *cycles) {
  double d[CHAINS][2];
  for (int k = 0; k < CHAINS; ++k) { d[k][0] = 0; d[k][1] = 0; }
  double a = 1e-3 * threadIdx.x, b = 1.0 + 1e-6 * threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) dmma(d[k][0], d[k][1], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int k = 0; k < CHAINS; ++k) s += d[k][0] + d[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}
// DFMA: one warp, CHAINS independent chains
template <int CHAINS>
__global__ void dfma_indep(double *out, int iters, long long *cycles) {
  double d[CHAINS];
  for (int k = 0; k < CHAINS; ++k) d[k] = k + 1e-3 * threadIdx.x;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) d[k] = fma(d[k], b, c);
  }
  long long t1 = clock64();
  double s = 0; for (int k = 0; k < CHAINS; ++k) s += d[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <int CH> void run_dmma(double *out, long long *cyc, int threads) {
  const int iters = 2048; long long c;
  dmma_indep<CH><<<1, threads>>>(out, iters, cyc); cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("  \"dmma_chains%d_warps_per_smsp%d_cyc_per_dmma_per_warp\": %.2f,\n", CH, threads / 128 ? threads / 128 : 1, (double)c / (iters * CH));
}
template <int CH> void run_dfma(double *out, long long *cyc, int threads) {
  const int iters = 2048; long long c;
  dfma_indep<CH><<<1, threads>>>(out, iters, cyc); cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("  \"dfma_chains%d_warps_per_smsp%d_cyc_per_dfma_per_warp\": %.2f,\n", CH, threads / 128 ? threads / 128 : 1, (double)c / (iters * CH));
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, sizeof(double) * 4096); cudaMalloc(&cyc, sizeof(long long));
  printf("{\n");
  for (int threads : {32, 128, 256, 512, 1024}) {
    run_dmma<1>(out, cyc, threads); run_dmma<2>(out, cyc, threads); run_dmma<4>(out, cyc, threads); run_dmma<8>(out, cyc, threads);
  }
  for (int threads : {32, 128, 256, 512}) { run_dfma<1>(out, cyc, threads); run_dfma<4>(out, cyc, threads); run_dfma<8>(out, cyc, threads); }
  printf("  \"note\": \"cycles per instruction per warp, block of N threads on one SM (warps spread over the 4 sub-partitions)\"\n}\n");
  return 0;
}
