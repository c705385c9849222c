// polish.cu -- OSQP's solution polishing (osqp/src/polish.c), both regimes.
// The reference leaves polish off (src/ModelPredictiveControlAPI.cpp:51-52 sets verbosity and warm start only); it is an
// opt-in setting here as in osqp-eigen (settings()->setPolish(true)): smpc_solver_set_polish.
//
// After the ADMM kernel, one warp per QP that ended SOLVED (persistent warps, grid-stride over the batch):
//   1. active set from the scaled iterates (form_Ared): lower-active  z_i - l_i < -y_i,  upper-active  u_i - z_i < y_i;
//   2. KKT system  [P̄ + delta I, Ared'; Ared, -delta I] [x; y_red] = [-q̄; l_low; u_upp]  assembled dense (lower triangle, row
//      stride N = n + #active) in the warp's workspace, factorised LDL' in place (quasi-definite: no pivoting), solved;
//   3. polish_refine_iter steps of iterative refinement against the unregularised KKT matrix;
//   4. z = A̅ x, y expanded, (z, y) projected onto the normal cone (proj.c project_normalcone), update_info on the polished
//      point; it replaces the ADMM solution only when it improves the residuals as polish.c tests (status_polish = 1),
//      otherwise the ADMM solution stays (status_polish = -1).
// QPs that did not end SOLVED get status_polish = 0.  The warp's workspace is a shared-memory slice while (n + m)^2 doubles
// of eight warps fit an SM, a slice of a global scratch buffer (L2 resident) otherwise.
#include "device_types.cuh"
#include "kernels.cuh"

namespace smpc {

namespace {
constexpr unsigned kFull = 0xffffffffu;
constexpr size_t kSmemBudget = 200 * 1024;
__device__ __forceinline__ double wmax_nn(double v) {   // max over the warp of non-negative doubles (IEEE order = integer order)
  const unsigned hi = (unsigned)__double2hiint(v);
  const unsigned mh = __reduce_max_sync(kFull, hi);
  const unsigned lo = hi == mh ? (unsigned)__double2loint(v) : 0u;
  const unsigned ml = __reduce_max_sync(kFull, lo);
  return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__host__ __device__ inline size_t polish_vec_doubles(int n, int m) {
  const size_t N = (size_t)n + m;
  return 5 * N + (size_t)n + 3 * (size_t)m + N;   // Dk b sol r t | q̄ | z y zp | act, sgn as ints in N doubles
}
}  // namespace

__global__ void __launch_bounds__(256) polish_kernel(PolishDataDev P, BatchDev Bt, SettingsDev S, double delta, int refine,
                                                     int *__restrict__ status_polish, double *__restrict__ scratch, int warps_per_cta) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = P.n, m = P.m, Nmax = n + m;
  const int gwarp = blockIdx.x * warps_per_cta + warp, nwarps = gridDim.x * warps_per_cta;
  const size_t vecs = polish_vec_doubles(n, m);
  double *K, *vbase;
  if (scratch) { K = scratch + (size_t)gwarp * Nmax * Nmax; vbase = smem + (size_t)warp * vecs; }
  else { K = smem + (size_t)warp * ((size_t)Nmax * Nmax + vecs); vbase = K + (size_t)Nmax * Nmax; }
  double *Dk = vbase, *bv = Dk + Nmax, *sol = bv + Nmax, *rr = sol + Nmax, *tt = rr + Nmax;
  double *qb = tt + Nmax, *zs = qb + n, *ys = zs + m, *zp = ys + m;
  int *act = reinterpret_cast<int *>(zp + m);
  const bool unscale = !S.scaled_termination;

  for (int b = gwarp; b < Bt.B; b += nwarps) {
    if (Bt.status[b] != SMPC_SOLVED) { if (lane == 0) status_polish[b] = 0; continue; }
    const double *Pb = P.Pbar + (size_t)b * P.strideP, *Ab = P.Abar + (size_t)b * P.strideA;
    const double *Dv = P.D + (size_t)b * P.strideD, *Ev = P.E + (size_t)b * P.strideE;
    const double c = P.c_inst ? P.c_inst[b] : P.c, cinv = 1.0 / c;
    __syncwarp();
    // ---- scaled data and iterates of this instance
    for (int i = lane; i < n; i += 32) qb[i] = Bt.q ? c * (Dv[i] * Bt.q[(size_t)b * n + i]) : 0.0;
    for (int r = lane; r < m; r += 32) { zs[r] = Bt.z[(size_t)b * m + r]; ys[r] = Bt.y[(size_t)b * m + r]; }
    __syncwarp();
    // ---- form_Ared: lower-active rows first, then upper-active rows (warp-ordered compaction keeps OSQP's row order)
    int k = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int r0 = 0; r0 < m; r0 += 32) {
        const int r = r0 + lane;
        bool on = false;
        double bound = 0.0;
        if (r < m) {
          const double lo = Ev[r] * (Bt.l ? Bt.l[(size_t)b * m + r] : P.l0[r]), hi = Ev[r] * (Bt.u ? Bt.u[(size_t)b * m + r] : P.u0[r]);
          on = pass == 0 ? (zs[r] - lo < -ys[r]) : (hi - zs[r] < ys[r]);
          bound = pass == 0 ? lo : hi;
        }
        const unsigned mask = __ballot_sync(kFull, on);
        if (on) { const int pos = k + __popc(mask & ((1u << lane) - 1u)); act[pos] = r; bv[n + pos] = bound; }
        k += __popc(mask);
      }
    const int N = n + k;
    for (int i = lane; i < n; i += 32) bv[i] = -qb[i];
    __syncwarp();
    // ---- K = [P̄ + delta I, Ared'; Ared, -delta I], lower triangle, row stride N
    for (int i = 0; i < N; ++i) {
      double *Ki = K + (size_t)i * N;
      if (i < n) { for (int j = lane; j <= i; j += 32) Ki[j] = Pb[(size_t)i * n + j] + (i == j ? delta : 0.0); }
      else {
        const double *arow = Ab + (size_t)act[i - n] * n;
        for (int j = lane; j <= i; j += 32) Ki[j] = j < n ? arow[j] : (j == i ? -delta : 0.0);
      }
    }
    __syncwarp();
    // ---- LDL' in place, right-looking: at step j lane-owned rows i > j subtract l_ij * (column j, still unscaled) from their
    //      trailing entries; column j is overwritten with L(:, j) once every row has used it.  Dk[j] keeps the pivot.
    bool ok = true;
    for (int j = 0; j < N; ++j) {
      const double d = K[(size_t)j * N + j];
      if (d == 0.0 || d != d) { ok = false; break; }
      if (lane == 0) Dk[j] = d;
      const double dinv = 1.0 / d;
      for (int i = j + 1 + lane; i < N; i += 32) tt[i] = K[(size_t)i * N + j];   // column j (unscaled), contiguous copy
      __syncwarp();
      for (int i = j + 1 + lane; i < N; i += 32) {
        const double lij = tt[i] * dinv;
        double *Ki = K + (size_t)i * N;
        for (int cc = j + 1; cc <= i; ++cc) Ki[cc] = fma(-lij, tt[cc], Ki[cc]);
        Ki[j] = lij;
      }
      __syncwarp();
    }
    int result = -1;
    if (ok) {
      auto kkt_solve = [&](double *v) {   // L D L' v = v in place; column-oriented substitutions
        for (int j = 0; j < N; ++j) {
          const double vj = v[j];
          for (int i = j + 1 + lane; i < N; i += 32) v[i] = fma(-K[(size_t)i * N + j], vj, v[i]);
          __syncwarp();
        }
        for (int i = lane; i < N; i += 32) v[i] /= Dk[i];
        __syncwarp();
        for (int j = N - 1; j > 0; --j) {
          const double vj = v[j];
          for (int i = lane; i < j; i += 32) v[i] = fma(-K[(size_t)j * N + i], vj, v[i]);
          __syncwarp();
        }
      };
      for (int i = lane; i < N; i += 32) sol[i] = bv[i];
      __syncwarp();
      kkt_solve(sol);
      for (int it = 0; it < refine; ++it) {
        // r = b - [P̄, Ared'; Ared, 0] sol
        for (int i = lane; i < N; i += 32) {
          double s = bv[i];
          if (i < n) {
            for (int j = 0; j < n; ++j) s = fma(-Pb[(size_t)i * n + j], sol[j], s);
            for (int a = 0; a < k; ++a) s = fma(-Ab[(size_t)act[a] * n + i], sol[n + a], s);
          } else {
            const double *arow = Ab + (size_t)act[i - n] * n;
            for (int j = 0; j < n; ++j) s = fma(-arow[j], sol[j], s);
          }
          rr[i] = s;
        }
        __syncwarp();
        kkt_solve(rr);
        for (int i = lane; i < N; i += 32) sol[i] += rr[i];
        __syncwarp();
      }
      // ---- polished point: x = sol[0:n]; y expanded; (z, y) <- normal-cone projection of (A̅ x, y); update_info
      for (int r = lane; r < m; r += 32) tt[r] = 0.0;          // y (expanded) in tt[0:m]   (Nmax >= m)
      __syncwarp();
      for (int a = lane; a < k; a += 32) tt[act[a]] = sol[n + a];
      __syncwarp();
      double u_rp = 0.0, s_rp = 0.0;
      for (int r = lane; r < m; r += 32) {
        const double *arow = Ab + (size_t)r * n;
        double s = 0.0;
        for (int j = 0; j < n; ++j) s = fma(arow[j], sol[j], s);
        const double lo = Ev[r] * (Bt.l ? Bt.l[(size_t)b * m + r] : P.l0[r]), hi = Ev[r] * (Bt.u ? Bt.u[(size_t)b * m + r] : P.u0[r]);
        const double zy = s + tt[r];
        const double zr = fmin(fmax(zy, lo), hi);
        zp[r] = zr; tt[r] = zy - zr;
        s_rp = fmax(s_rp, fabs(s - zr)); u_rp = fmax(u_rp, fabs((1.0 / Ev[r]) * (s - zr)));
      }
      __syncwarp();
      double u_rd = 0.0, s_rd = 0.0, ob = 0.0;
      for (int i = lane; i < n; i += 32) {
        double px = 0.0, aty = 0.0;
        for (int j = 0; j < n; ++j) px = fma(Pb[(size_t)i * n + j], sol[j], px);
        for (int r = 0; r < m; ++r) aty = fma(Ab[(size_t)r * n + i], tt[r], aty);
        const double rd = (qb[i] + px) + aty;
        s_rd = fmax(s_rd, fabs(rd)); u_rd = fmax(u_rd, fabs((1.0 / Dv[i]) * rd));
        ob += 0.5 * sol[i] * px + qb[i] * sol[i];
      }
      const double pri = m == 0 ? 0.0 : (unscale ? wmax_nn(u_rp) : wmax_nn(s_rp));
      const double dua = unscale ? cinv * wmax_nn(u_rd) : wmax_nn(s_rd);
      ob = wsum(ob);
      const double pri0 = Bt.pri_res[b], dua0 = Bt.dua_res[b];
      const bool success = (pri < pri0 && dua < dua0) || (pri < pri0 && dua0 < 1e-10) || (dua < dua0 && pri0 < 1e-10);
      if (success && pri == pri && dua == dua) {
        result = 1;
        for (int i = lane; i < n; i += 32) {
          Bt.x_out[(size_t)b * n + i] = Dv[i] * sol[i];
          if (P.VinvT == nullptr) Bt.xi[(size_t)b * n + i] = sol[i];
          else { double s = 0.0; for (int j = 0; j < n; ++j) s = fma(P.VinvT[(size_t)j * n + i], sol[j], s); Bt.xi[(size_t)b * n + i] = s; }   // xi = V^-1 x̄
        }
        for (int r = lane; r < m; r += 32) {
          if (Bt.y_out) Bt.y_out[(size_t)b * m + r] = cinv * (Ev[r] * tt[r]);
          Bt.z[(size_t)b * m + r] = zp[r]; Bt.y[(size_t)b * m + r] = tt[r];
        }
        if (lane == 0) { Bt.obj[b] = cinv * ob; Bt.pri_res[b] = pri; Bt.dua_res[b] = dua; }
      }
    }
    if (lane == 0) status_polish[b] = result;
  }
}

size_t polish_scratch_doubles(int n, int m, int num_sms) {
  const size_t N = (size_t)n + m;
  const size_t per = (N * N + polish_vec_doubles(n, m)) * sizeof(double);
  return per <= kSmemBudget ? 0 : (size_t)num_sms * 8 * N * N;
}

cudaError_t launch_polish(const PolishDataDev &P, const BatchDev &Bt, const SettingsDev &S, double delta, int refine,
                          int *status_polish, double *scratch, int num_sms, cudaStream_t stream) {
  const size_t N = (size_t)P.n + P.m;
  const size_t vec = polish_vec_doubles(P.n, P.m) * sizeof(double), per = N * N * sizeof(double) + vec;
  int wpc = 8, grid;
  size_t smem;
  if (per <= kSmemBudget) {
    while (wpc > 1 && wpc * per > kSmemBudget) --wpc;
    smem = wpc * per; scratch = nullptr;
    grid = (Bt.B + wpc - 1) / wpc;
    if (grid > 4 * num_sms) grid = 4 * num_sms;
  } else {
    if (!scratch || wpc * vec > kSmemBudget) return cudaErrorInvalidValue;
    smem = wpc * vec;
    grid = num_sms;
  }
  cudaError_t e = cudaFuncSetAttribute(polish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device: every launch
  if (e != cudaSuccess) return e;
  polish_kernel<<<grid, wpc * 32, smem, stream>>>(P, Bt, S, delta, refine, status_polish, scratch, wpc);
  return cudaGetLastError();
}

}  // namespace smpc
