// ilp2_iteration.cu -- would a warp that carries TWO small QPs through one instruction stream beat one QP per warp?
// The body is the row-pair iteration of admm_shared_small_kernel<true> (16 + 8 DFMA, two shared-memory broadcasts, three
// 64-bit shuffles, the element-wise z / y / w update) on synthetic bounded data, no termination checks.  The shared
// operators (m1, wr) are held once per lane, the iterates once per QP.  Reported: instance-iterations per second of the
// whole GPU for Q QPs per warp and C resident CTAs (of four warps) per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ilp2_iteration ilp2_iteration.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr unsigned kFull = 0xffffffffu;
// the kernel's own volatile accesses.  (They also keep the compiler from interleaving the Q streams of a warp: a round of Q = 2
// costs exactly twice one QP.  Non-volatile asm loads are no test of that -- the compiler hoists them out of the loop.)
__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// FOLD = true: rho and 1 / (1 + rho lambda) folded into the register operators (two DMUL fewer per iteration)
template <int Q, int CTAS, bool FOLD>
__global__ void __launch_bounds__(128, CTAS) body(double *out, int iters) {
  __shared__ __align__(16) double sm[4 * Q * 48];                 // per warp and QP: [xi; wd] (32) + t (16)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> 4, i = lane & 15;
  double m1[16], wr[8];
  for (int j = 0; j < 16; ++j) m1[j] = 1e-2 * ((j * 7 + lane * 3) % 11 - 5);
  for (int j = 0; j < 8; ++j) wr[j] = 1e-2 * ((j * 5 + lane) % 9 - 4);
  double xi[Q], om_xi[Q], base[Q], nqh[Q], dinv[Q], rv[Q], lb[Q], ub[Q];
  uint32_t a_cv[Q], a_t[Q], a_tz[Q], a_xi[Q], a_w[Q];
  for (int q = 0; q < Q; ++q) {
    double *cb = sm + (warp * Q + q) * 48;
    for (int e = lane; e < 48; e += 32) cb[e] = 1e-3 * (e % 7);
    xi[q] = 0.01 * q; om_xi[q] = 0.0; base[q] = 0.0; nqh[q] = h ? 0.0 : 0.05 * (i % 3 - 1); dinv[q] = 0.7; rv[q] = 0.1;
    lb[q] = -0.5; ub[q] = 0.5;
    a_cv[q] = (uint32_t)__cvta_generic_to_shared(cb + 16 * h);
    a_t[q] = (uint32_t)__cvta_generic_to_shared(cb + 32 + i);
    a_tz[q] = (uint32_t)__cvta_generic_to_shared(cb + 32 + 8 * h);
    a_xi[q] = (uint32_t)__cvta_generic_to_shared(cb + i);
    a_w[q] = (uint32_t)__cvta_generic_to_shared(cb + 16 + i);
  }
  const double alpha = 1.6, oma = -0.6, alpha_r = h ? -alpha : alpha;
  __syncwarp();
  for (int s = 0; s < iters; ++s) {
    double t[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      double a0 = nqh[q], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double2 u0 = lds128(a_cv[q] + 32 * j), u1 = lds128(a_cv[q] + 32 * j + 16);
        a0 = fma(m1[4 * j + 0], u0.x, a0); a1 = fma(m1[4 * j + 1], u0.y, a1);
        a2 = fma(m1[4 * j + 2], u1.x, a2); a3 = fma(m1[4 * j + 3], u1.y, a3);
      }
      double acc = (a0 + a1) + (a2 + a3);
      acc += __shfl_xor_sync(kFull, acc, 16);
      t[q] = FOLD ? acc : acc * dinv[q];
      if (h == 0) sts64(a_t[q], t[q]);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const double xn = fma(alpha, t[q], om_xi[q]);
      xi[q] = xn; om_xi[q] = oma * xn;
      if (h == 0) sts64(a_xi[q], xn);
      double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double2 u0 = lds128(a_tz[q] + 32 * j), u1 = lds128(a_tz[q] + 32 * j + 16);
        b0 = fma(wr[4 * j + 0], u0.x, b0); b1 = fma(wr[4 * j + 1], u0.y, b1);
        b2 = fma(wr[4 * j + 2], u1.x, b2); b3 = fma(wr[4 * j + 3], u1.y, b3);
      }
      double zt = (b0 + b1) + (b2 + b3);
      zt += __shfl_xor_sync(kFull, zt, 16);
      const double v = fma(alpha_r, zt, base[q]);
      const double zn = v < lb[q] ? lb[q] : (v > ub[q] ? ub[q] : v);
      const double dn = v - zn;
      const double w = FOLD ? fma(2.0, zn, -v) : rv[q] * fma(2.0, zn, -v);
      const double wo = __shfl_xor_sync(kFull, w, 16);
      if (h == 0) sts64(a_w[q], w - wo);
      base[q] = fma(oma, zn, dn);
    }
    __syncwarp();
  }
  double acc = 0.0;
  for (int q = 0; q < Q; ++q) acc += xi[q] + base[q];
  out[blockIdx.x * 128 + threadIdx.x] = acc;
}

// Two QPs per warp with the two instruction streams interleaved BY HAND (loads of both, then the arithmetic of both, then
// the shuffles and stores of both), so that the dependent chains overlap even with volatile shared-memory accesses.
template <int CTAS>
__global__ void __launch_bounds__(128, CTAS) body2(double *out, int iters) {
  constexpr int Q = 2;
  __shared__ __align__(16) double sm[4 * Q * 48];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> 4, i = lane & 15;
  double m1[16], wr[8];
  for (int j = 0; j < 16; ++j) m1[j] = 1e-2 * ((j * 7 + lane * 3) % 11 - 5);
  for (int j = 0; j < 8; ++j) wr[j] = 1e-2 * ((j * 5 + lane) % 9 - 4);
  double xi[Q], om_xi[Q], base[Q], nqh[Q], dinv[Q], rv[Q], lb[Q], ub[Q];
  uint32_t a_cv[Q], a_t[Q], a_tz[Q], a_xi[Q], a_w[Q];
  for (int q = 0; q < Q; ++q) {
    double *cb = sm + (warp * Q + q) * 48;
    for (int e = lane; e < 48; e += 32) cb[e] = 1e-3 * (e % 7);
    xi[q] = 0.01 * q; om_xi[q] = 0.0; base[q] = 0.0; nqh[q] = h ? 0.0 : 0.05 * (i % 3 - 1); dinv[q] = 0.7; rv[q] = 0.1;
    lb[q] = -0.5; ub[q] = 0.5;
    a_cv[q] = (uint32_t)__cvta_generic_to_shared(cb + 16 * h);
    a_t[q] = (uint32_t)__cvta_generic_to_shared(cb + 32 + i);
    a_tz[q] = (uint32_t)__cvta_generic_to_shared(cb + 32 + 8 * h);
    a_xi[q] = (uint32_t)__cvta_generic_to_shared(cb + i);
    a_w[q] = (uint32_t)__cvta_generic_to_shared(cb + 16 + i);
  }
  const double alpha = 1.6, oma = -0.6, alpha_r = h ? -alpha : alpha;
  __syncwarp();
  for (int s = 0; s < iters; ++s) {
    double2 c[Q][8];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) c[q][j] = lds128(a_cv[q] + 16 * j);
    double acc[Q], t[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      double a0 = nqh[q], a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a0 = fma(m1[4 * j + 0], c[q][2 * j].x, a0); a1 = fma(m1[4 * j + 1], c[q][2 * j].y, a1);
        a2 = fma(m1[4 * j + 2], c[q][2 * j + 1].x, a2); a3 = fma(m1[4 * j + 3], c[q][2 * j + 1].y, a3);
      }
      acc[q] = (a0 + a1) + (a2 + a3);
    }
    double sh[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) sh[q] = __shfl_xor_sync(kFull, acc[q], 16);
#pragma unroll
    for (int q = 0; q < Q; ++q) t[q] = (acc[q] + sh[q]) * dinv[q];
#pragma unroll
    for (int q = 0; q < Q; ++q)
      if (h == 0) sts64(a_t[q], t[q]);
    __syncwarp();
    double2 tt[Q][4];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) tt[q][j] = lds128(a_tz[q] + 16 * j);
    double zt[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const double xn = fma(alpha, t[q], om_xi[q]);
      xi[q] = xn; om_xi[q] = oma * xn;
      double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        b0 = fma(wr[4 * j + 0], tt[q][2 * j].x, b0); b1 = fma(wr[4 * j + 1], tt[q][2 * j].y, b1);
        b2 = fma(wr[4 * j + 2], tt[q][2 * j + 1].x, b2); b3 = fma(wr[4 * j + 3], tt[q][2 * j + 1].y, b3);
      }
      zt[q] = (b0 + b1) + (b2 + b3);
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) sh[q] = __shfl_xor_sync(kFull, zt[q], 16);
    double w[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const double v = fma(alpha_r, zt[q] + sh[q], base[q]);
      const double zn = v < lb[q] ? lb[q] : (v > ub[q] ? ub[q] : v);
      const double dn = v - zn;
      w[q] = rv[q] * fma(2.0, zn, -v);
      base[q] = fma(oma, zn, dn);
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) sh[q] = __shfl_xor_sync(kFull, w[q], 16);
#pragma unroll
    for (int q = 0; q < Q; ++q)
      if (h == 0) { sts64(a_xi[q], xi[q]); sts64(a_w[q], w[q] - sh[q]); }
    __syncwarp();
  }
  double r = 0.0;
  for (int q = 0; q < Q; ++q) r += xi[q] + base[q];
  out[blockIdx.x * 128 + threadIdx.x] = r;
}

template <int CTAS>
static void run2(int sms, int iters, double *out) {
  const int grid = sms * CTAS;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  body2<CTAS><<<grid, 128>>>(out, iters);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    body2<CTAS><<<grid, 128>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, body2<CTAS>);
  const double qps = (double)grid * 4 * 2, rate = qps * iters / (best * 1e-3);
  printf("{\"qps_per_warp\": 2, \"hand_interleaved\": 1, \"ctas_per_sm\": %d, \"registers\": %d, \"ms\": %.4f, \"cycles_per_iteration_round\": %.1f, "
         "\"instance_iterations_per_s\": %.4e}\n", CTAS, fa.numRegs, best, best * 1e-3 * 1.965e9 / iters, rate);
}

template <int Q, int CTAS, bool FOLD = false>
static void run(int sms, int iters, double *out) {
  const int grid = sms * CTAS;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  body<Q, CTAS, FOLD><<<grid, 128>>>(out, iters);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    body<Q, CTAS, FOLD><<<grid, 128>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, body<Q, CTAS, FOLD>);
  const double qps = (double)grid * 4 * Q, rate = qps * iters / (best * 1e-3);
  printf("{\"qps_per_warp\": %d, \"ctas_per_sm\": %d, \"folded_scalings\": %d, \"registers\": %d, \"ms\": %.4f, \"cycles_per_iteration_round\": %.1f, "
         "\"instance_iterations_per_s\": %.4e}\n", Q, CTAS, (int)FOLD, fa.numRegs, best, best * 1e-3 * 1.965e9 / iters, rate);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * 148 * 4 * 128 * 2);
  const int iters = 4000;
  run<1, 1>(sms, iters, out);
  run<1, 2>(sms, iters, out);
  run<1, 3>(sms, iters, out);
  run<1, 4>(sms, iters, out);
  run<2, 1>(sms, iters, out);
  run<2, 2>(sms, iters, out);
  run<2, 3>(sms, iters, out);
  run<3, 2>(sms, iters, out);
  run<4, 2>(sms, iters, out);
  run<1, 3, true>(sms, iters, out);
  run2<1>(sms, iters, out);
  run2<2>(sms, iters, out);
  run2<3>(sms, iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
