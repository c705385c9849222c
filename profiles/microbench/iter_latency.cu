// iter_latency.cu -- where do the cycles of one DMMA-tile ADMM iteration go?  One CTA of 4 warps runs the main-loop body of
// admm_shared_small_mma_kernel (no events) with parts removed, timed with clock64 over many iterations.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o iter_latency iter_latency.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int OFF> __device__ __forceinline__ double lds64o(uint32_t addr) {
  double v; asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF) : "memory"); return v;
}
template <int OFF> __device__ __forceinline__ void sts64o(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(v) : "memory");
}
template <int OFF> __device__ __forceinline__ void sts128o(uint32_t addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(addr), "n"(OFF), "d"(a), "d"(b) : "memory");
}
constexpr int NP = 16, MP = 32;
constexpr int oS = 0, oP0 = oS + (NP + MP) * 8, oP1 = oP0 + NP * 8, oEnd = oP1 + NP * 8;

// MODE bits: 1 = GEMM 1, 2 = barrier A, 4 = t phase, 8 = GEMM 2 DMMAs, 16 = epilogue, 32 = barrier B
template <int MODE>
__global__ void __launch_bounds__(128) body(double *out, int iters, long long *cycles) {
  __shared__ __align__(16) double pan[oEnd];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q2 = 2 * (lane & 3), bfrag = (lane & 3) * 8 + g;
  for (int e = tid; e < oEnd; e += 128) pan[e] = 1e-3 * (e % 7);
  const int rb1 = warp & 1, kh = warp >> 1;
  double a1[6], a2[4];
  for (int j = 0; j < 6; ++j) a1[j] = 1e-2 * (j + lane % 5);
  for (int j = 0; j < 4; ++j) a2[j] = 1e-2 * (j + lane % 3);
  const uint32_t sPan = (uint32_t)__cvta_generic_to_shared(pan);
  const uint32_t aB = sPan + 8 * bfrag, aB1 = aB + 8 * 192 * kh, aPartSt = sPan + 8 * ((kh ? oP1 : oP0) + (8 * rb1 + g) * 8 + q2);
  const uint32_t aXi = aB + 8 * 32 * warp, aC = sPan + 8 * ((8 * warp + g) * 8 + q2);
  const bool sel1 = warp & 1, sel2 = warp & 2;
  double r_nq[2] = {0.1, 0.2}, r_dv[4] = {0.5, 0.4, 0.3, 0.2}, r_xi = 0, r_omxi = 0, r_y[2] = {0, 0}, r_lo[2] = {-1, -1}, r_hi[2] = {1, 1};
  double r_rv[2] = {0.1, 0.1}, r_ri[2] = {10, 10}, r_base[2] = {0, 0};
  const double alpha = 1.6, oma = -0.6;
  __syncthreads();
  const long long t0 = clock64();
  double t[4] = {0.1, 0.2, 0.3, 0.4};
  for (int k = 0; k < iters; ++k) {
    if (MODE & 1) {
      const double b0 = lds64o<0>(aB1), b1 = lds64o<256>(aB1), b2 = lds64o<512>(aB1), b3 = lds64o<768>(aB1), b4 = lds64o<1024>(aB1), b5 = lds64o<1280>(aB1);
      double c0[2] = {r_nq[0], r_nq[1]}, c1[2] = {0.0, 0.0};
      dmma(c0, a1[0], b0); dmma(c1, a1[1], b1); dmma(c0, a1[2], b2); dmma(c1, a1[3], b3); dmma(c0, a1[4], b4); dmma(c1, a1[5], b5);
      sts128o<0>(aPartSt, c0[0] + c1[0], c0[1] + c1[1]);
    }
    if (MODE & 2) __syncthreads();
    if (MODE & 4) {
      t[0] = (lds64o<8 * oP0>(aB) + lds64o<8 * oP1>(aB)) * r_dv[0];
      t[1] = (lds64o<8 * (oP0 + 32)>(aB) + lds64o<8 * (oP1 + 32)>(aB)) * r_dv[1];
      t[2] = (lds64o<8 * (oP0 + 64)>(aB) + lds64o<8 * (oP1 + 64)>(aB)) * r_dv[2];
      t[3] = (lds64o<8 * (oP0 + 96)>(aB) + lds64o<8 * (oP1 + 96)>(aB)) * r_dv[3];
    }
    double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
    if (MODE & 8) { dmma(c0, a2[0], t[0]); dmma(c1, a2[1], t[1]); dmma(c0, a2[2], t[2]); dmma(c1, a2[3], t[3]); }
    else { c0[0] = t[0]; c0[1] = t[1]; c1[0] = t[2]; c1[1] = t[3]; }
    if (MODE & 16) {
      const double ta = sel1 ? t[1] : t[0], tb = sel1 ? t[3] : t[2], tw = sel2 ? tb : ta;
      const double xn = fma(alpha, tw, r_omxi);
      sts64o<8 * oS>(aXi, xn);
      r_xi = xn; r_omxi = oma * xn;
      double wv[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double v = fma(alpha, c0[j] + c1[j], r_base[j]);
        const double zn = v < r_lo[j] ? r_lo[j] : (v > r_hi[j] ? r_hi[j] : v);
        const double yn = r_rv[j] * (v - zn);
        wv[j] = r_rv[j] * fma(2.0, zn, -v);
        r_y[j] = yn;
        r_base[j] = fma(r_ri[j], yn, oma * zn);
      }
      sts128o<8 * (oS + NP * 8)>(aC, wv[0], wv[1]);
    } else {
      t[0] += c0[0] * 1e-9; t[1] += c0[1] * 1e-9; t[2] += c1[0] * 1e-9; t[3] += c1[1] * 1e-9;
    }
    if (MODE & 32) __syncthreads();
  }
  const long long t1 = clock64();
  out[tid] = t[0] + t[1] + t[2] + t[3] + r_y[0] + r_y[1] + r_xi + r_base[0] + r_base[1];
  if (tid == 0) *cycles = t1 - t0;
}

template <int MODE> void run(const char *name, double *out, long long *cyc) {
  const int iters = 4096; long long c;
  body<MODE><<<1, 128>>>(out, iters, cyc); cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  printf("  \"%s\": %.1f,\n", name, (double)c / iters);
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, sizeof(double) * 128); cudaMalloc(&cyc, sizeof(long long));
  printf("{\n");
  run<63>("full_iteration_cycles", out, cyc);
  run<63 - 2 - 32>("no_barriers", out, cyc);
  run<1 + 2>("gemm1_plus_barrier", out, cyc);
  run<1>("gemm1_only", out, cyc);
  run<4>("t_phase_only", out, cyc);
  run<4 + 8>("t_phase_plus_gemm2_dmma", out, cyc);
  run<4 + 8 + 16>("t_phase_gemm2_epilogue", out, cyc);
  run<4 + 8 + 16 + 32>("t_phase_gemm2_epilogue_barrier", out, cyc);
  run<2 + 32>("two_barriers_only", out, cyc);
  printf("  \"note\": \"cycles per iteration, one CTA of 4 warps alone on an SM\"\n}\n");
  return 0;
}
