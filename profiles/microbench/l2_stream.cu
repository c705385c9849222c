// Microbenchmark: L2 -> SM read throughput of the tile kernel's operator stream on B200.
// Every CTA (one per SM, 16 warps) reads the SAME L2-resident buffer (the packed operator, 320 KB for n = 200) again and again with the
// kernel's own access: one ld.global.nc.L1::no_allocate.v2.f64 per lane = 512 contiguous bytes per warp, `depth` loads in flight per
// warp.  Prints bytes / cycle / SM and the chip-wide TB/s for several depths and buffer sizes: the ceiling admm_shared_tile_kernel's
// GEMMs run against (bench.py roofline.l2_operator_stream).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l2_stream l2_stream.cu && ./l2_stream
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 ldg_stream(const double2 *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

template <int DEPTH>
__global__ void __launch_bounds__(512, 1) stream_kernel(const double2 *buf, int chunks, int passes, double *sink, long long *cycles, int skew) {
  // chunks = 512-byte warp chunks in the buffer; warp w of the CTA reads chunks w, w + 16, ... (the row-blocks of a GEMM)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int p = 0; p < passes; ++p) {
    // skew: CTA j starts j * skew chunks into the buffer (tiles of the real kernel drift apart: their requests for a line do not coincide)
    for (int c0 = warp; c0 < chunks; c0 += 16 * DEPTH) {
      const int c = (c0 + (int)blockIdx.x * skew) % chunks;
      double2 v[DEPTH];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) { const int cc = (c + 16 * d) % chunks; v[d] = ldg_stream(buf + (size_t)cc * 32 + lane); }
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) acc += v[d].x + v[d].y;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456) sink[0] = acc;
}

template <int DEPTH>
void run(const double2 *buf, size_t bytes, int sms, double *sink, long long *cyc, int skew) {
  const int chunks = (int)(bytes / 512), passes = (int)((64u << 20) / bytes) + 1;
  stream_kernel<DEPTH><<<sms, 512>>>(buf, chunks, 2, sink, cyc, skew);   // warm the L2
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  stream_kernel<DEPTH><<<sms, 512>>>(buf, chunks, passes, sink, cyc, skew);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double mean = 0; long long mx = 0; for (int i = 0; i < sms; ++i) { mean += h[i]; if (h[i] > mx) mx = h[i]; } mean /= sms;
  const double per_sm = (double)chunks * 512.0 * passes;
  printf("{\"buffer_kb\": %zu, \"cta_skew_chunks\": %d, \"loads_in_flight_per_warp\": %d, \"bytes_per_cycle_per_sm\": %.2f, \"bytes_per_cycle_per_sm_slowest\": %.2f, \"chip_tb_s\": %.2f}\n",
         bytes >> 10, skew, DEPTH, per_sm / mean, per_sm / (double)mx, per_sm * sms / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double2 *buf; double *sink; long long *cyc;
  cudaMalloc(&buf, 8u << 20); cudaMemset(buf, 0, 8u << 20); cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8 * 256);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d}\n", prop.name, sms, khz / 1000);
  for (int skew : {0, 7, 37})
    for (size_t kb : {320u, 640u}) {
      run<1>(buf, kb << 10, sms, sink, cyc, skew);
      run<3>(buf, kb << 10, sms, sink, cyc, skew);
      run<6>(buf, kb << 10, sms, sink, cyc, skew);
    }
  return 0;
}
