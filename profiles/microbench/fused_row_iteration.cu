// fused_row_iteration.cu -- one shared-memory round trip per ADMM iteration instead of two.
// The row-pair iteration of admm_shared_small_kernel<true> is  t = dinv .* (sigma G xi + W_top' wd - q̂),  z̃_top = W_top t,
// i.e. two dependent mat-vecs, each ending in a half-warp shuffle and a shared-memory broadcast (392 cycles alone on an SM).
// Substituting t gives ONE 32 x 32 mat-vec on the state s = [xi; wd]:
//     [t; z̃_top] = K(rho) s - k0(rho),   K = [dinv .* M1; W_top diag(dinv) M1],  M1 = [sigma G | W_top']
// with one row of K(rho) per lane (32 registers of operator, per instance because rho is), no reduction shuffle, and the
// element-wise update of BOTH rows of a pair on the lane that holds z̃_top(i) (no exchange shuffle either):
// 1024 instead of 768 MACs per instance-iteration, but a dependent chain of one broadcast + 8 DFMA + ~7 element-wise ops.
// Reported like ilp2_iteration.cu: cycles per iteration and instance-iterations/s of the whole GPU for C four-warp CTAs per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fused_row_iteration fused_row_iteration.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 lds128(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

// CHAINS = number of independent accumulators of the 32-term row product
template <int CTAS, int CHAINS>
__global__ void __launch_bounds__(128, CTAS) body(double *out, int iters) {
  __shared__ __align__(16) double sm[4 * 32];                 // per warp: s = [xi (16); wd (16)]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, h = lane >> 4, i = lane & 15;
  double kr[32];
  for (int j = 0; j < 32; ++j) kr[j] = 1e-2 * ((j * 7 + lane * 3) % 11 - 5);
  double *cb = sm + warp * 32;
  cb[lane] = 1e-3 * (lane % 7);
  const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(cb), a_me = (uint32_t)__cvta_generic_to_shared(cb + lane);
  double k0 = 0.05 * (i % 3 - 1);
  double xi = 0.01, om_xi = 0.0, base_t = 0.0, base_b = 0.0, rv = 0.1;
  const double lb_t = -0.5, ub_t = 0.5, lb_b = -0.4, ub_b = 0.6;
  const double alpha = 1.6, oma = -0.6;
  __syncwarp();
  for (int s = 0; s < iters; ++s) {
    double a[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) a[c] = c == 0 ? k0 : 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const double2 u = lds128(a_s + 16 * j);
      a[(2 * j) % CHAINS] = fma(kr[2 * j], u.x, a[(2 * j) % CHAINS]);
      a[(2 * j + 1) % CHAINS] = fma(kr[2 * j + 1], u.y, a[(2 * j + 1) % CHAINS]);
    }
    double acc;
    if (CHAINS == 4) acc = (a[0] + a[1]) + (a[2] + a[3]);
    else if (CHAINS == 2) acc = a[0] + a[1];
    else { acc = 0.0; for (int c = 0; c < CHAINS; ++c) acc += a[c]; }
    __syncwarp();                                             // every lane has read s
    if (h == 0) {
      const double xn = fma(alpha, acc, om_xi);
      xi = xn; om_xi = oma * xn;
      sts64(a_me, xn);
    } else {
      const double vt = fma(alpha, acc, base_t), vb = fma(-alpha, acc, base_b);
      const double zt = vt < lb_t ? lb_t : (vt > ub_t ? ub_t : vt);
      const double zb = vb < lb_b ? lb_b : (vb > ub_b ? ub_b : vb);
      const double dt = vt - zt, db = vb - zb;
      const double wt = rv * fma(2.0, zt, -vt), wb = rv * fma(2.0, zb, -vb);
      sts64(a_me, wt - wb);
      base_t = fma(oma, zt, dt); base_b = fma(oma, zb, db);
    }
    __syncwarp();
  }
  out[blockIdx.x * 128 + threadIdx.x] = xi + base_t + base_b;
}


// 2-D blocks: lane (a, b) = (lane >> 2, lane & 3) holds a 4 x 8 block of K(rho) (rows of group a, columns 8b .. 8b+7), reads only
// ITS 8 entries of s (four LDS.128 with four distinct addresses per instruction instead of sixteen full broadcasts: a quarter of
// the shared-memory -> register traffic), and the four partial sums are transpose-reduced over the four b-lanes with three
// 64-bit shuffles (the local row order of every lane is permuted in the pack so that no select is needed).  The element-wise
// code is uniform: a xi-lane is a degenerate pair (bounds +-inf, rv = (1, 0), alpha_b = 0).
template <int CTAS>
__global__ void __launch_bounds__(128, CTAS) body2d(double *out, int iters) {
  __shared__ __align__(16) double sm[4 * 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = lane & 3;
  double kr[4][8];
  for (int l = 0; l < 4; ++l)
    for (int j = 0; j < 8; ++j) kr[l][j] = 1e-2 * ((j * 7 + lane * 3 + l * 5) % 11 - 5);
  double *cb = sm + warp * 32;
  cb[lane] = 1e-3 * (lane % 7);
  // s is stored in 16-byte chunks ordered (j, b): the four distinct chunks one LDS.128 reads are contiguous (no bank conflict)
  const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(cb) + 16 * b;
  const uint32_t a_me = (uint32_t)__cvta_generic_to_shared(cb) + 8 * (((((lane & 7) >> 1) * 4 + (lane >> 3)) * 2) + (lane & 1));
  const bool is_xi = b < 2;
  const double k0 = 0.05 * (lane % 3 - 1);
  double base_t = 0.0, base_b = 0.0;
  const double rv_t = is_xi ? 1.0 : 0.1, rv_b = is_xi ? 0.0 : 0.1;
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const double lb_t = is_xi ? -inf : -0.5, ub_t = is_xi ? inf : 0.5, lb_b = -0.4, ub_b = 0.6;
  const double alpha = 1.6, oma = -0.6, alpha_b = is_xi ? 0.0 : -alpha;
  __syncwarp();
  for (int s = 0; s < iters; ++s) {
    double2 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = lds128(a_s + 64 * j);
    double p[4];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      double e0 = l == 0 ? k0 : 0.0, e1 = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) { e0 = fma(kr[l][2 * j], u[j].x, e0); e1 = fma(kr[l][2 * j + 1], u[j].y, e1); }
      p[l] = e0 + e1;
    }
    const double q0 = p[0] + __shfl_xor_sync(0xffffffffu, p[2], 1), q1 = p[1] + __shfl_xor_sync(0xffffffffu, p[3], 1);
    const double acc = q0 + __shfl_xor_sync(0xffffffffu, q1, 2);
    const double vt = fma(alpha, acc, base_t), vb = fma(alpha_b, acc, base_b);
    const double zt = vt < lb_t ? lb_t : (vt > ub_t ? ub_t : vt);
    const double zb = vb < lb_b ? lb_b : (vb > ub_b ? ub_b : vb);
    const double dt = vt - zt, db = vb - zb;
    sts64(a_me, fma(rv_t, fma(2.0, zt, -vt), -(rv_b * fma(2.0, zb, -vb))));
    base_t = fma(oma, zt, dt); base_b = fma(oma, zb, db);
    __syncwarp();
  }
  out[blockIdx.x * 128 + threadIdx.x] = base_t + base_b;
}

template <int CTAS>
static void run2d(int sms, int iters, double *out) {
  const int grid = sms * CTAS;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  body2d<CTAS><<<grid, 128>>>(out, iters);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    body2d<CTAS><<<grid, 128>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, body2d<CTAS>);
  const double qps = (double)grid * 4, rate = qps * iters / (best * 1e-3);
  printf("{\"layout\": \"2d_blocks\", \"ctas_per_sm\": %d, \"registers\": %d, \"ms\": %.4f, \"cycles_per_iteration\": %.1f, "
         "\"instance_iterations_per_s\": %.4e}\n", CTAS, fa.numRegs, best, best * 1e-3 * 1.965e9 / iters, rate);
}

template <int CTAS, int CHAINS>
static void run(int sms, int iters, double *out) {
  const int grid = sms * CTAS;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  body<CTAS, CHAINS><<<grid, 128>>>(out, iters);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0);
    body<CTAS, CHAINS><<<grid, 128>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, body<CTAS, CHAINS>);
  const double qps = (double)grid * 4, rate = qps * iters / (best * 1e-3);
  printf("{\"ctas_per_sm\": %d, \"chains\": %d, \"registers\": %d, \"ms\": %.4f, \"cycles_per_iteration\": %.1f, "
         "\"instance_iterations_per_s\": %.4e}\n", CTAS, CHAINS, fa.numRegs, best, best * 1e-3 * 1.965e9 / iters, rate);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * 148 * 4 * 128);
  const int iters = 4000;
  run<1, 4>(sms, iters, out);
  run<2, 4>(sms, iters, out);
  run<3, 4>(sms, iters, out);
  run<4, 4>(sms, iters, out);
  run<1, 8>(sms, iters, out);
  run<3, 8>(sms, iters, out);
  run<1, 2>(sms, iters, out);
  run<3, 2>(sms, iters, out);
  run2d<1>(sms, iters, out);
  run2d<2>(sms, iters, out);
  run2d<3>(sms, iters, out);
  run2d<4>(sms, iters, out);
  run2d<5>(sms, iters, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
