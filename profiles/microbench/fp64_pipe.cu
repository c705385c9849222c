// fp64_pipe.cu -- micro-benchmark of the B200 FP64 paths the ADMM kernels can use:
//   DFMA  throughput (independent chains, full occupancy) and dependent-chain latency,
//   DMMA  (mma.sync.aligned.m8n8k4.f64) throughput and dependent latency,
//   smem broadcast LDS.128 and 64-bit SHFL rates.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipe fp64_pipe.cu ; prints one JSON line.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_tput(double *out, int iters) {
  double a[8]; 
  for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-3 + k;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
  }
  double s = 0; for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dfma_lat(double *out, int iters, long long *cycles) {
  double a = threadIdx.x * 1e-3; const double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a = fma(a, b, c);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void dmma_tput(double *out, int iters) {
  double d[8][2];
  for (int k = 0; k < 8; ++k) { d[k][0] = 0; d[k][1] = 0; }
  double a = 1e-3 * threadIdx.x, b = 1.0 + 1e-6 * threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) dmma(d[k][0], d[k][1], a, b);
  }
  double s = 0; for (int k = 0; k < 8; ++k) s += d[k][0] + d[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma_lat(double *out, int iters, long long *cycles) {
  double d0 = 0, d1 = 0, a = 1e-3 * threadIdx.x, b = 1.0 + 1e-6 * threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) dmma(d0, d1, a, b);
  }
  long long t1 = clock64();
  out[threadIdx.x] = d0 + d1;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}
__global__ void lds_shfl(double *out, int iters, long long *cycles) {
  __shared__ __align__(16) double buf[64];
  if (threadIdx.x < 64) buf[threadIdx.x] = threadIdx.x;
  __syncthreads();
  double s = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) { double2 v = reinterpret_cast<const double2 *>(buf)[(k + i) & 31]; s += v.x + v.y; }
  }
  long long t1 = clock64();
  double r = s;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) r += __shfl_xor_sync(0xffffffffu, r, 16);
  }
  long long t2 = clock64();
  out[threadIdx.x] = r;
  if (threadIdx.x == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t1; }
}

static float time_kernel(void (*launch)(void)) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
static double *g_out; static long long *g_cyc;
static const int kIters = 4096, kGrid = 148 * 8, kBlock = 256;
static void l_dfma() { dfma_tput<<<kGrid, kBlock>>>(g_out, kIters); }
static void l_dmma() { dmma_tput<<<kGrid, kBlock>>>(g_out, kIters); }

int main() {
  cudaMalloc(&g_out, sizeof(double) * kGrid * kBlock); cudaMalloc(&g_cyc, 4 * sizeof(long long));
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  float ms_f = time_kernel(l_dfma), ms_m = time_kernel(l_dmma);
  double dfma_tflops = 2.0 * kGrid * kBlock * 8.0 * kIters / (ms_f * 1e-3) / 1e12;
  double dmma_tflops = 2.0 * (kGrid * (kBlock / 32.0)) * 8.0 * kIters * 256.0 / (ms_m * 1e-3) / 1e12;
  long long c[4];
  dfma_lat<<<1, 32>>>(g_out, 1024, g_cyc); cudaMemcpy(c, g_cyc, sizeof(long long), cudaMemcpyDeviceToHost);
  double lat_dfma = (double)c[0] / (1024 * 16);
  dmma_lat<<<1, 32>>>(g_out, 1024, g_cyc); cudaMemcpy(c, g_cyc, sizeof(long long), cudaMemcpyDeviceToHost);
  double lat_dmma = (double)c[0] / (1024 * 16);
  lds_shfl<<<1, 32>>>(g_out, 1024, g_cyc); cudaMemcpy(c, g_cyc, 2 * sizeof(long long), cudaMemcpyDeviceToHost);
  // one warp alone: issue-bound cost per LDS.128 broadcast (+2 DADD) and per dependent 64-bit SHFL (+DADD)
  printf("{\"device\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_tflops\": %.3f, \"dfma_dep_latency_cyc\": %.2f, "
         "\"dmma_dep_latency_cyc\": %.2f, \"lds128_bcast_plus_2dadd_cyc\": %.2f, \"shfl64_dep_plus_dadd_cyc\": %.2f}\n",
         p.name, p.multiProcessorCount, dfma_tflops, dmma_tflops, lat_dfma, lat_dmma, (double)c[0] / (1024 * 16), (double)c[1] / (1024 * 16));
  return 0;
}
