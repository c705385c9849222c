// admm_instance.cu -- per-instance regime: every QP carries its OWN P_i and A_i (per-instance linearised
// plants, BASELINE config 4), so nothing of the KKT system is shared.  One warp owns one QP for its whole solve:
//   setup  (ruiz_instance_kernel): OSQP scale_data (10 Ruiz passes + cost scaling) on (P_i, A_i) in shared memory
//   solve  (admm_instance_kernel): A̅_i and the factor live in the warp's shared-memory slice;
//          M = P̄ + sigma I + A̅' diag(rho_vec) A̅  ->  Cholesky  ->  M^-1 = L^-T L^-1   (refactored whenever rho changes,
//          exactly when OSQP refactors its KKT matrix),  then the OSQP iteration in x-space:
//          x̃ = M^-1 (sigma x - q̄ + A̅'(rho_vec.*z - y)),  z̃ = A̅ x̃,  relaxation, clip to [l̄, ū], dual update,
//          residual norms / termination / infeasibility / rho adaptation every check interval.
// Restates OSQP 0.6.x -- the arithmetic the reference reaches through solver.initSolver() /
// solver.solve() (src/ModelPredictiveControlAPI.cpp:64,102).  Shared-memory rows are padded to an odd number of
// doubles so that "lane = row" and "lane = column" sweeps are both bank-conflict free.
#include <cstdio>

#include "device_types.cuh"
#include "kernels.cuh"

namespace smpc {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr double kMaxScaling = 1e4;

// exact max over the warp of NON-NEGATIVE doubles (every use below reduces absolute values): compare the IEEE bit
// patterns as two 32-bit halves with REDUX instead of five 64-bit shuffle + compare steps
__device__ __forceinline__ double wmax(double v) {
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
__device__ __forceinline__ double limit_scaling(double v) {
  v = v < kMinScaling ? 1.0 : v;
  return v > kMaxScaling ? kMaxScaling : v;
}
__host__ __device__ __forceinline__ int odd_ld(int n) { return n | 1; }
// per-warp shared-memory doubles of admm_instance_kernel: A̅ (m x ld), M^-1 and scratch (n x ld each), 8 n-vectors, 10 m-vectors
__host__ __device__ __forceinline__ size_t instance_warp_doubles(int n, int m) {
  return (size_t)m * odd_ld(n) + 2 * (size_t)n * odd_ld(n) + 8 * (size_t)n + 10 * (size_t)m;
}

// sum_r A[r*ld + i] * vec[r]   (lane = column i)
__device__ __forceinline__ double col_dot_s(const double *A, int ld, int rows, int i, const double *vec) {
  double a0 = 0.0, a1 = 0.0;
  int r = 0;
  for (; r + 2 <= rows; r += 2) { a0 = fma(A[r * ld + i], vec[r], a0); a1 = fma(A[(r + 1) * ld + i], vec[r + 1], a1); }
  if (r < rows) a0 = fma(A[r * ld + i], vec[r], a0);
  return a0 + a1;
}
// sum_k A[r*ld + k] * vec[k]   (lane = row r)
__device__ __forceinline__ double row_dot_s(const double *A, int ld, int cols, int r, const double *vec) {
  double a0 = 0.0, a1 = 0.0;
  int k = 0;
  for (; k + 2 <= cols; k += 2) { a0 = fma(A[r * ld + k], vec[k], a0); a1 = fma(A[r * ld + k + 1], vec[k + 1], a1); }
  if (k < cols) a0 = fma(A[r * ld + k], vec[k], a0);
  return a0 + a1;
}

struct Info {
  double pri_res, dua_res, nEz, nEAx, nDq, nDAty, nDPx;
  double s_rp, s_rd, s_z, s_Ax, s_q, s_Aty, s_Px;
  double obj;
};

}  // namespace

// ---------------------------------------------------------------------------------------------- setup
// In-place OSQP scale_data on instance b: P[b] (n*n, upper triangle mirrored first), A[b] (m*n); writes D, E, c.
__global__ void ruiz_instance_kernel(InstanceDataDev I, int iters, int warps_per_cta) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * warps_per_cta + warp;
  if (b >= I.B) return;
  const int n = I.n, m = I.m, ldP = odd_ld(n), ldA = odd_ld(n);
  double *P = smem + (size_t)warp * ((n + m) * ldP + 2 * n + 2 * m);
  double *A = P + n * ldP, *D = A + m * ldA, *dt = D + n, *E = dt + n, *et = E + m;
  double *gP = I.P + (size_t)b * n * n, *gA = I.A + (size_t)b * m * n;
  for (int e = lane; e < n * n; e += 32) { int i = e / n, j = e % n; P[i * ldP + j] = i <= j ? gP[i * n + j] : gP[j * n + i]; }
  for (int e = lane; e < m * n; e += 32) A[(e / n) * ldA + e % n] = gA[e];
  for (int j = lane; j < n; j += 32) D[j] = 1.0;
  for (int i = lane; i < m; i += 32) E[i] = 1.0;
  double c = 1.0;
  __syncwarp();
  for (int it = 0; it < iters; ++it) {
    for (int j = lane; j < n; j += 32) {
      double r = 0.0;
      for (int i = 0; i < n; ++i) r = fmax(r, fabs(P[i * ldP + j]));
      for (int i = 0; i < m; ++i) r = fmax(r, fabs(A[i * ldA + j]));
      dt[j] = 1.0 / sqrt(limit_scaling(r));
    }
    for (int i = lane; i < m; i += 32) {
      double r = 0.0;
      for (int j = 0; j < n; ++j) r = fmax(r, fabs(A[i * ldA + j]));
      et[i] = 1.0 / sqrt(limit_scaling(r));
    }
    __syncwarp();
    for (int e = lane; e < n * n; e += 32) { int i = e / n, j = e % n; P[i * ldP + j] = (dt[i] * P[i * ldP + j]) * dt[j]; }
    for (int e = lane; e < m * n; e += 32) { int i = e / n, j = e % n; A[i * ldA + j] = (et[i] * A[i * ldA + j]) * dt[j]; }
    for (int j = lane; j < n; j += 32) D[j] *= dt[j];
    for (int i = lane; i < m; i += 32) E[i] *= et[i];
    __syncwarp();
    // cost scaling: mean column norm of P̄ (the setup gradient is 0: its norm counts as 1)
    for (int j = lane; j < n; j += 32) {
      double r = 0.0;
      for (int i = 0; i < n; ++i) r = fmax(r, fabs(P[i * ldP + j]));
      dt[j] = r;
    }
    __syncwarp();
    double mean = 0.0;
    for (int j = 0; j < n; ++j) mean += dt[j];   // same summation order as the oracle
    mean /= n;
    double ct = limit_scaling(fmax(mean, 1.0));
    ct = 1.0 / ct;
    __syncwarp();
    for (int e = lane; e < n * n; e += 32) { int i = e / n, j = e % n; P[i * ldP + j] *= ct; }
    c *= ct;
    __syncwarp();
  }
  for (int e = lane; e < n * n; e += 32) gP[e] = P[(e / n) * ldP + e % n];
  for (int e = lane; e < m * n; e += 32) gA[e] = A[(e / n) * ldA + e % n];
  for (int j = lane; j < n; j += 32) I.D[(size_t)b * n + j] = D[j];
  for (int i = lane; i < m; i += 32) I.E[(size_t)b * m + i] = E[i];
  if (lane == 0) I.c[b] = c;
}

// ---------------------------------------------------------------------------------------------- solve
__global__ void __launch_bounds__(256) admm_instance_kernel(InstanceDataDev I, BatchDev Bt, SettingsDev S, int warps_per_cta) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * warps_per_cta + warp;
  if (b >= Bt.B) return;
  const int n = I.n, m = I.m, ldA = odd_ld(n), ldM = odd_ld(n);
  double *base = smem + (size_t)warp * instance_warp_doubles(n, m);
  double *Ab = base, *Mi = Ab + m * ldA, *Sc = Mi + n * ldM;
  double *x = Sc + n * ldM, *dx = x + n, *qb = dx + n, *xt = qb + n, *sPx = xt + n, *sAty = sPx + n, *Dv = sAty + n, *Dinv = Dv + n;
  double *z = Dinv + n, *y = z + m, *lb = y + m, *ub = lb + m, *w = ub + m, *zt = w + m, *dy = zt + m, *rv = dy + m, *Ev = rv + m, *Einv = Ev + m;
  const double *gP = I.P + (size_t)b * n * n, *gA = I.A + (size_t)b * m * n;
  const double alpha = S.alpha, c = I.c[b], cinv = 1.0 / c;
  const bool unscale = !S.scaled_termination;
  const bool warm = S.warm_start && !Bt.fresh;

  // ---- load the instance
  for (int e = lane; e < m * n; e += 32) Ab[(e / n) * ldA + e % n] = gA[e];
  for (int i = lane; i < n; i += 32) {
    const double d = I.D[(size_t)b * n + i];
    Dv[i] = d; Dinv[i] = 1.0 / d;
    qb[i] = Bt.q ? c * (d * Bt.q[(size_t)b * n + i]) : 0.0;
    x[i] = warm ? Bt.xi[(size_t)b * n + i] : 0.0;
    dx[i] = 0.0;
  }
  double rho = Bt.fresh ? fmin(fmax(S.rho0, kRhoMin), kRhoMax) : Bt.rho[b];
  int bad_rows = 0;
  for (int r = lane; r < m; r += 32) {
    const double e = I.E[(size_t)b * m + r];
    Ev[r] = e; Einv[r] = 1.0 / e;
    lb[r] = e * (Bt.l ? Bt.l[(size_t)b * m + r] : I.l0[r]);
    ub[r] = e * (Bt.u ? Bt.u[(size_t)b * m + r] : I.u0[r]);
    z[r] = warm ? Bt.z[(size_t)b * m + r] : 0.0;
    y[r] = warm ? Bt.y[(size_t)b * m + r] : 0.0;
    dy[r] = 0.0;
    bad_rows |= lb[r] > ub[r];
  }
  const bool bad_bounds = __any_sync(kFull, bad_rows);
  int rho_updates = 0;
  __syncwarp();

  // rho_vec from the instance's own (scaled) bounds: OSQP set_rho_vec / update_rho_vec
  auto set_rho_vec = [&]() {
    for (int r = lane; r < m; r += 32) {
      const bool fr = lb[r] < -kInfty * kMinScaling && ub[r] > kInfty * kMinScaling;
      rv[r] = fr ? kRhoMin : ((ub[r] - lb[r] < kRhoTolRow) ? kRhoEqOverIneq * rho : rho);
    }
    __syncwarp();
  };

  // M = P̄ + sigma I + A̅' diag(rho_vec) A̅ ;  M = L L' ;  Mi = M^-1 = L^-T L^-1.   Returns false if M is not PD.
  auto refactor = [&]() -> bool {
    for (int e = lane; e < n * n; e += 32) {
      const int i = e / n, j = e % n;
      if (j > i) continue;
      double s0 = gP[i * n + j] + (i == j ? S.sigma : 0.0), s1 = 0.0;
      int r = 0;
      for (; r + 2 <= m; r += 2) {
        s0 = fma(rv[r] * Ab[r * ldA + i], Ab[r * ldA + j], s0);
        s1 = fma(rv[r + 1] * Ab[(r + 1) * ldA + i], Ab[(r + 1) * ldA + j], s1);
      }
      if (r < m) s0 = fma(rv[r] * Ab[r * ldA + i], Ab[r * ldA + j], s0);
      Sc[i * ldM + j] = s0 + s1;
    }
    __syncwarp();
    // left-looking Cholesky, lower triangle of Sc in place
    int ok = 1;
    for (int j = 0; j < n; ++j) {
      for (int i = j + lane; i < n; i += 32) {
        double s = Sc[i * ldM + j];
        for (int k = 0; k < j; ++k) s = fma(-Sc[i * ldM + k], Sc[j * ldM + k], s);
        xt[i] = s;    // column j before scaling (xt is free here)
      }
      __syncwarp();
      const double d = xt[j];
      if (!(d > 0.0)) { ok = 0; break; }
      const double sd = sqrt(d);
      for (int i = j + lane; i < n; i += 32) Sc[i * ldM + j] = i == j ? sd : xt[i] / sd;
      __syncwarp();
    }
    if (!ok) return false;
    // Linv (lower) into Mi: lane c solves L v = e_c by forward substitution, column by column
    for (int cc = lane; cc < n; cc += 32) {
      Mi[cc * ldM + cc] = 1.0 / Sc[cc * ldM + cc];
      for (int i = cc + 1; i < n; ++i) {
        double s = 0.0;
        for (int k = cc; k < i; ++k) s = fma(Sc[i * ldM + k], Mi[k * ldM + cc], s);
        Mi[i * ldM + cc] = -s / Sc[i * ldM + i];
      }
    }
    __syncwarp();
    // M^-1 = Linv' Linv (full symmetric) into Sc
    for (int e = lane; e < n * n; e += 32) {
      const int i = e / n, j = e % n;
      if (j > i) continue;
      double s = 0.0;
      for (int k = i; k < n; ++k) s = fma(Mi[k * ldM + i], Mi[k * ldM + j], s);
      Sc[i * ldM + j] = s;
    }
    __syncwarp();
    for (int e = lane; e < n * n; e += 32) { const int i = e / n, j = e % n; if (j > i) Sc[i * ldM + j] = Sc[j * ldM + i]; }
    __syncwarp();
    double *t = Mi; Mi = Sc; Sc = t;
    return true;
  };

  int status = SMPC_UNSOLVED, iter = 0;
  bool can_check = false;
  Info F = {};

  auto update_info = [&]() {
    for (int i = lane; i < n; i += 32) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s = fma(gP[i * n + k], x[k], s);   // P̄ x (P̄ symmetric, read from L2 at checks only)
      sPx[i] = s;
      sAty[i] = col_dot_s(Ab, ldA, m, i, y);
    }
    for (int r = lane; r < m; r += 32) w[r] = row_dot_s(Ab, ldA, n, r, x);
    __syncwarp();
    double a_rp = 0, a_z = 0, a_Ax = 0, u_rp = 0, u_z = 0, u_Ax = 0;
    for (int r = lane; r < m; r += 32) {
      const double rp = w[r] - z[r], ei = Einv[r];
      a_rp = fmax(a_rp, fabs(rp)); a_z = fmax(a_z, fabs(z[r])); a_Ax = fmax(a_Ax, fabs(w[r]));
      u_rp = fmax(u_rp, fabs(ei * rp)); u_z = fmax(u_z, fabs(ei * z[r])); u_Ax = fmax(u_Ax, fabs(ei * w[r]));
    }
    double a_rd = 0, a_q = 0, a_Aty = 0, a_Px = 0, u_rd = 0, u_q = 0, u_Aty = 0, u_Px = 0, ob = 0;
    for (int i = lane; i < n; i += 32) {
      const double rd = (qb[i] + sPx[i]) + sAty[i], di = Dinv[i];
      a_rd = fmax(a_rd, fabs(rd)); a_q = fmax(a_q, fabs(qb[i])); a_Aty = fmax(a_Aty, fabs(sAty[i])); a_Px = fmax(a_Px, fabs(sPx[i]));
      u_rd = fmax(u_rd, fabs(di * rd)); u_q = fmax(u_q, fabs(di * qb[i])); u_Aty = fmax(u_Aty, fabs(di * sAty[i])); u_Px = fmax(u_Px, fabs(di * sPx[i]));
      ob += 0.5 * x[i] * sPx[i] + qb[i] * x[i];
    }
    F.s_rp = wmax(a_rp); F.s_z = wmax(a_z); F.s_Ax = wmax(a_Ax);
    F.s_rd = wmax(a_rd); F.s_q = wmax(a_q); F.s_Aty = wmax(a_Aty); F.s_Px = wmax(a_Px);
    if (unscale) {
      F.pri_res = wmax(u_rp); F.nEz = wmax(u_z); F.nEAx = wmax(u_Ax);
      F.dua_res = cinv * wmax(u_rd); F.nDq = wmax(u_q); F.nDAty = wmax(u_Aty); F.nDPx = wmax(u_Px);
      F.obj = cinv * wsum(ob);
    } else {
      F.pri_res = F.s_rp; F.nEz = F.s_z; F.nEAx = F.s_Ax;
      F.dua_res = F.s_rd; F.nDq = F.s_q; F.nDAty = F.s_Aty; F.nDPx = F.s_Px;
      F.obj = wsum(ob);
    }
    if (m == 0) F.pri_res = 0.0;
  };

  auto primal_infeasible = [&](double eps) -> bool {
    double nd = 0.0;
    for (int r = lane; r < m; r += 32) {
      double d = dy[r];
      const bool uinf = ub[r] > kInfty * kMinScaling, linf = lb[r] < -kInfty * kMinScaling;
      if (uinf) d = linf ? 0.0 : fmin(d, 0.0); else if (linf) d = fmax(d, 0.0);
      dy[r] = d;
      nd = fmax(nd, fabs(unscale ? Ev[r] * d : d));
    }
    nd = wmax(nd);
    __syncwarp();
    if (!(nd > eps)) return false;
    double lhs = 0.0;
    for (int r = lane; r < m; r += 32) {
      const double dp = fmax(dy[r], 0.0), dm = fmin(dy[r], 0.0);
      if (dp != 0.0) lhs += ub[r] * dp;
      if (dm != 0.0) lhs += lb[r] * dm;
    }
    lhs = wsum(lhs);
    if (!(lhs < -eps * nd)) return false;
    double na = 0.0;
    for (int i = lane; i < n; i += 32) {
      const double v = col_dot_s(Ab, ldA, m, i, dy);
      na = fmax(na, fabs(unscale ? Dinv[i] * v : v));
    }
    return wmax(na) < eps * nd;
  };

  auto dual_infeasible = [&](double eps) -> bool {
    double nd = 0.0, qd = 0.0;
    for (int i = lane; i < n; i += 32) {
      nd = fmax(nd, fabs(unscale ? Dv[i] * dx[i] : dx[i]));
      qd += qb[i] * dx[i];
    }
    nd = wmax(nd); qd = wsum(qd);
    const double cs = unscale ? c : 1.0;
    if (!(nd > eps)) return false;
    if (!(qd < -cs * eps * nd)) return false;
    double np = 0.0;
    for (int i = lane; i < n; i += 32) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s = fma(gP[i * n + k], dx[k], s);
      np = fmax(np, fabs(unscale ? Dinv[i] * s : s));
    }
    if (!(wmax(np) < cs * eps * nd)) return false;
    int bad = 0;
    for (int r = lane; r < m; r += 32) {
      double v = row_dot_s(Ab, ldA, n, r, dx);
      if (unscale) v *= Einv[r];
      if (((ub[r] < kInfty * kMinScaling) && (v > eps * nd)) || ((lb[r] > -kInfty * kMinScaling) && (v < -eps * nd))) bad = 1;
    }
    return !__any_sync(kFull, bad);
  };

  auto check_termination = [&](bool approx) -> bool {
    double ea = S.eps_abs, er = S.eps_rel, epi = S.eps_prim_inf, edi = S.eps_dual_inf;
    if (approx) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (m == 0) prim_ok = true;
    else if (F.pri_res < ea + er * fmax(F.nEz, F.nEAx)) prim_ok = true;
    else prim_inf = primal_infeasible(epi);
    if (F.dua_res < ea + er * (unscale ? cinv : 1.0) * fmax(fmax(F.nDq, F.nDAty), F.nDPx)) dual_ok = true;
    else dual_inf = dual_infeasible(edi);
    if (prim_ok && dual_ok) { status = approx ? SMPC_SOLVED_INACCURATE : SMPC_SOLVED; return true; }
    if (prim_inf) { status = approx ? SMPC_PRIMAL_INFEASIBLE_INACCURATE : SMPC_PRIMAL_INFEASIBLE; F.obj = kInfty; return true; }
    if (dual_inf) { status = approx ? SMPC_DUAL_INFEASIBLE_INACCURATE : SMPC_DUAL_INFEASIBLE; F.obj = -kInfty; return true; }
    return false;
  };

  bool factor_ok = true;
  if (!bad_bounds) {
    set_rho_vec();
    factor_ok = refactor();
  }
  for (iter = 1; iter <= S.max_iter && !bad_bounds && factor_ok; ++iter) {
    for (int r = lane; r < m; r += 32) w[r] = rv[r] * z[r] - y[r];
    __syncwarp();
    for (int i = lane; i < n; i += 32) sAty[i] = (S.sigma * x[i] - qb[i]) + col_dot_s(Ab, ldA, m, i, w);   // rhs
    __syncwarp();
    for (int i = lane; i < n; i += 32) xt[i] = row_dot_s(Mi, ldM, n, i, sAty);                              // x̃ = M^-1 rhs
    __syncwarp();
    for (int r = lane; r < m; r += 32) {
      const double ztl = row_dot_s(Ab, ldA, n, r, xt);
      const double rr = rv[r], rinv = 1.0 / rr;
      const double zr = alpha * ztl + (1.0 - alpha) * z[r];
      const double zn = fmin(fmax(zr + rinv * y[r], lb[r]), ub[r]);
      const double d = rr * (zr - zn);
      z[r] = zn; y[r] += d; dy[r] = d;
    }
    for (int i = lane; i < n; i += 32) {
      const double xn = alpha * xt[i] + (1.0 - alpha) * x[i];
      dx[i] = xn - x[i];
      x[i] = xn;
    }
    __syncwarp();
    can_check = S.check_every && (iter % S.check_every == 0);
    if (can_check) {
      update_info();
      if (check_termination(false)) break;
    }
    if (S.adaptive_rho && S.rho_interval && (iter % S.rho_interval == 0)) {
      if (!can_check) update_info();
      const double pr = F.s_rp / (fmax(F.s_z, F.s_Ax) + kDivTol);
      const double dr = F.s_rd / (fmax(fmax(F.s_q, F.s_Aty), F.s_Px) + kDivTol);
      const double rn = fmin(fmax(rho * sqrt(pr / (dr + kDivTol)), kRhoMin), kRhoMax);
      if (rn > rho * S.rho_tol || rn < rho / S.rho_tol) {
        rho = rn; ++rho_updates;
        __syncwarp();
        set_rho_vec();
        if (!refactor()) { factor_ok = false; break; }
      }
    }
    __syncwarp();
  }
  if (iter > S.max_iter) iter = S.max_iter;
  if (bad_bounds || !factor_ok) iter = 0;
  else {
    if (!can_check) { update_info(); check_termination(false); }
    if (status == SMPC_UNSOLVED) { if (!check_termination(true)) status = SMPC_MAX_ITER_REACHED; }
  }

  const bool has_sol = !bad_bounds && factor_ok && !(status == SMPC_PRIMAL_INFEASIBLE || status == SMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                                                     status == SMPC_DUAL_INFEASIBLE || status == SMPC_DUAL_INFEASIBLE_INACCURATE);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    if (Bt.x_out) Bt.x_out[(size_t)b * n + i] = has_sol ? Dv[i] * x[i] : qnan;
    Bt.xi[(size_t)b * n + i] = has_sol ? x[i] : 0.0;
  }
  for (int r = lane; r < m; r += 32) {
    if (Bt.y_out) Bt.y_out[(size_t)b * m + r] = has_sol ? cinv * (Ev[r] * y[r]) : qnan;
    Bt.z[(size_t)b * m + r] = has_sol ? z[r] : 0.0;
    Bt.y[(size_t)b * m + r] = has_sol ? y[r] : 0.0;
  }
  if (lane == 0) {
    Bt.rho[b] = rho;
    Bt.status[b] = status; Bt.iter[b] = iter; Bt.rho_updates[b] = rho_updates;
    Bt.obj[b] = F.obj; Bt.pri_res[b] = F.pri_res; Bt.dua_res[b] = F.dua_res;
  }
}

// --------------------------------------------------------------------------------------- solve, n <= 32, m <= 64
// Register-operator variant of admm_instance_kernel (same algorithm, same order of the OSQP steps): the rows of A̅ that a
// lane multiplies in z̃ = A̅ x̃ (rows lane and lane + 32) and its row of M^-1 live in REGISTERS, A̅' is kept row-major in
// shared memory (row i = column i of A̅) so that A̅'w is a row sweep too, and every dot product reads its operands with
// 16-byte shared-memory loads into four independent accumulators.  Per iteration and lane: 160 DFMA, 96 LDS.128
// (v1: 150 DFMA behind 300 dependent LDS.64 in two accumulators).  Rows of A̅' are padded to ldT = 14 (mod 16) doubles:
// 16-byte row sweeps by 8 consecutive lanes then hit 8 different 16-byte bank groups.
constexpr int kRegN = 32, kRegM = 64;
__host__ __device__ __forceinline__ int reg_ldT(int m) { int l = (m + 1) & ~1; while ((l & 15) != 14) l += 2; return l; }
__host__ __device__ __forceinline__ size_t instance_reg_warp_doubles(int n, int m, bool paired) {
  return (size_t)n * reg_ldT(paired ? m / 2 : m) + (size_t)n * reg_ldT(n) + 8 * (size_t)kRegN + 11 * (size_t)kRegM;
}
// sum_{k < len} row[k] * vec[k], len even, both 16-byte aligned
__device__ __forceinline__ double dot_s128(const double *row, const double *vec, int len) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int k = 0;
#pragma unroll 4
  for (; k + 4 <= len; k += 4) {
    const double2 r0 = *reinterpret_cast<const double2 *>(row + k), v0 = *reinterpret_cast<const double2 *>(vec + k);
    const double2 r1 = *reinterpret_cast<const double2 *>(row + k + 2), v1 = *reinterpret_cast<const double2 *>(vec + k + 2);
    a0 = fma(r0.x, v0.x, a0); a1 = fma(r0.y, v0.y, a1); a2 = fma(r1.x, v1.x, a2); a3 = fma(r1.y, v1.y, a3);
  }
  if (k < len) {
    const double2 r0 = *reinterpret_cast<const double2 *>(row + k), v0 = *reinterpret_cast<const double2 *>(vec + k);
    a0 = fma(r0.x, v0.x, a0); a1 = fma(r0.y, v0.y, a1);
  }
  return (a0 + a1) + (a2 + a3);
}
// sum_{k < 32} reg[k] * vec[k] with the row in registers (zero padded) and vec (zero padded to 32) broadcast from shared memory
__device__ __forceinline__ double dot_reg32(const double (&reg)[kRegN], const double *vec) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
  for (int k = 0; k < kRegN; k += 4) {
    const double2 v0 = *reinterpret_cast<const double2 *>(vec + k), v1 = *reinterpret_cast<const double2 *>(vec + k + 2);
    a0 = fma(reg[k], v0.x, a0); a1 = fma(reg[k + 1], v0.y, a1); a2 = fma(reg[k + 2], v1.x, a2); a3 = fma(reg[k + 3], v1.y, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

// prepare != 0 (create time): no solve; stores S0, T (M(rho) = S0 + rho T for the setup bounds' row classes) and the rows of
// M(rho0)^-1, so that a solve starts without a factorisation (OSQP factors at osqp_setup, not in osqp_solve) and a rho update
// re-forms M with n(n+1)/2 FMAs instead of n(n+1)/2 length-m dot products.
// PAIRED: every instance's rows come as [G; -G] (the reference's two-sided limit, cpp:335; checked on the scaled data at
// create time): only G (mp = m / 2 rows) is kept -- G' in shared memory, one row of G per lane in registers --
// A̅'v = G'(v_top - v_bot) and (A̅ v)_bot = -(A̅ v)_top: 90 instead of 160 DFMA per lane-iteration, half the operator footprint.
template <bool PAIRED>
__global__ void __launch_bounds__(256, 1) admm_instance_reg_kernel(InstanceDataDev I, BatchDev Bt, SettingsDev S, int warps_per_cta, int prepare) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * warps_per_cta + warp;
  if (b >= Bt.B) return;
  const int n = I.n, m = I.m, mp = PAIRED ? m / 2 : m;                  // rows of the operator kept on chip
  const int ldT = reg_ldT(mp), ldM = reg_ldT(n), mv = (mp + 1) & ~1;
  double *base = smem + (size_t)warp * ((instance_reg_warp_doubles(n, m, PAIRED) + 1) & ~(size_t)1);
  // vectors first (strides kRegN / kRegM, zero padded, 16-byte aligned), then A̅' and the two n x n scratch matrices
  double *x = base, *dx = x + kRegN, *qb = dx + kRegN, *xt = qb + kRegN, *sPx = xt + kRegN, *sAty = sPx + kRegN, *Dv = sAty + kRegN, *Dinv = Dv + kRegN;
  double *z = Dinv + kRegN, *y = z + kRegM, *lb = y + kRegM, *ub = lb + kRegM, *w = ub + kRegM, *zt = w + kRegM, *dy = zt + kRegM, *rv = dy + kRegM, *Ev = rv + kRegM, *Einv = Ev + kRegM;
  double *vd = Einv + kRegM;                       // PAIRED scratch: v_top - v_bot (mp entries) / top half of A̅ v
  double *At = vd + kRegM, *Sc = At + n * ldT;
  const double *gP = I.P + (size_t)b * n * n, *gA = I.A + (size_t)b * m * n;
  const int li = lane < n ? lane : n - 1;          // lanes >= n sweep a valid row and drop the result
#ifdef SMPC_PROFILE
  long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pc = clock64();
#define PFI(i) { const long long tt = clock64(); pf[i] += tt - pc; pc = tt; }
#else
#define PFI(i)
#endif
  for (int e = lane; e < 8 * kRegN + 11 * kRegM; e += 32) base[e] = 0.0;
  for (int e = lane; e < n * ldT; e += 32) At[e] = 0.0;
  __syncwarp();
  const double alpha = S.alpha, c = I.c[b], cinv = 1.0 / c;
  const bool unscale = !S.scaled_termination;
  const bool warm = S.warm_start && !Bt.fresh && !prepare;

  // ---- load the instance: A̅' (PAIRED: G') into shared memory, rows lane (and lane + 32) of it into registers
  for (int row = 0; row < mp; ++row)                // coalesced row reads, transposed (conflict-free) shared-memory writes
    for (int col = lane; col < n; col += 32) At[col * ldT + row] = gA[row * n + col];
  __syncwarp();
  double ar0[kRegN], ar1[PAIRED ? 1 : kRegN], mi[kRegN];
#pragma unroll
  for (int k = 0; k < kRegN; ++k) {
    ar0[k] = (k < n && lane < mp) ? At[k * ldT + lane] : 0.0;
    if constexpr (!PAIRED) ar1[k] = (k < n && lane + 32 < mp) ? At[k * ldT + lane + 32] : 0.0;
    mi[k] = 0.0;
  }
  // (A̅' v)_li for an m-vector v in shared memory (valid on lanes < n)
  auto AT_dot = [&](const double *v) -> double {
    if (PAIRED) {
      __syncwarp();
      if (lane < mp) vd[lane] = v[lane] - v[lane + mp];
      __syncwarp();
      return dot_s128(At + li * ldT, vd, mv);
    }
    return dot_s128(At + li * ldT, v, mv);
  };
  // out[r] = (A̅ v)_r for every row r < m, v an n-vector in shared memory (zero padded to 32)
  auto A_rows = [&](const double *v, double *out) {
    if (PAIRED) {
      const double top = dot_reg32(ar0, v);
      if (lane < mp) { out[lane] = top; out[lane + mp] = -top; }
    } else if constexpr (!PAIRED) {
      if (lane < m) out[lane] = dot_reg32(ar0, v);
      if (lane + 32 < m) out[lane + 32] = dot_reg32(ar1, v);
    }
    __syncwarp();
  };
  for (int i = lane; i < n; i += 32) {
    const double d = I.D[(size_t)b * n + i];
    Dv[i] = d; Dinv[i] = 1.0 / d;
    qb[i] = Bt.q ? c * (d * Bt.q[(size_t)b * n + i]) : 0.0;
    x[i] = warm ? Bt.xi[(size_t)b * n + i] : 0.0;
    dx[i] = 0.0;
  }
  double rho = (Bt.fresh || prepare) ? fmin(fmax(S.rho0, kRhoMin), kRhoMax) : Bt.rho[b];
  int bad_rows = 0;
  for (int r = lane; r < m; r += 32) {
    const double e = I.E[(size_t)b * m + r];
    Ev[r] = e; Einv[r] = 1.0 / e;
    lb[r] = e * (Bt.l ? Bt.l[(size_t)b * m + r] : I.l0[r]);
    ub[r] = e * (Bt.u ? Bt.u[(size_t)b * m + r] : I.u0[r]);
    z[r] = warm ? Bt.z[(size_t)b * m + r] : 0.0;
    y[r] = warm ? Bt.y[(size_t)b * m + r] : 0.0;
    dy[r] = 0.0;
    bad_rows |= lb[r] > ub[r];
  }
  const bool bad_bounds = __any_sync(kFull, bad_rows);
  int rho_updates = 0;
  __syncwarp();

  // rho_vec from the instance's own (scaled) bounds: OSQP set_rho_vec / update_rho_vec
  // row class: -1 free, 0 inequality, 1 equality (OSQP constr_type)
  auto row_class = [&](double lo, double hi) { return (lo < -kInfty * kMinScaling && hi > kInfty * kMinScaling) ? -1 : ((hi - lo < kRhoTolRow) ? 1 : 0); };
  bool split_ok = I.S0 != nullptr;   // the prepared S0 / T / Minv0 hold for the row classes of the SETUP bounds only
  auto set_rho_vec = [&]() {
    int differs = 0;
    for (int r = lane; r < m; r += 32) {
      const int ct = row_class(lb[r], ub[r]);
      rv[r] = ct < 0 ? kRhoMin : (ct == 1 ? kRhoEqOverIneq * rho : rho);
      differs |= ct != row_class(Ev[r] * I.l0[r], Ev[r] * I.u0[r]);
    }
    if (__any_sync(kFull, differs)) split_ok = false;
    __syncwarp();
  };

  // M = P̄ + sigma I + A̅' diag(rho_vec) A̅ (lower triangle, in Sc) ;  M = L L' in place ;  then lane c solves
  // L L' v = e_c with v in its registers: v is row c of M^-1, exactly the operand of x̃ = M^-1 rhs.  No explicit L^-1,
  // no M^-1 in shared memory.  Returns false if M is not positive definite.
  auto refactor = [&]() -> bool {
    PFI(2)
    const int tri = n * (n + 1) / 2;
    for (int e = lane; e < tri; e += 32) {
      int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
      while ((i + 1) * (i + 2) / 2 <= e) ++i;
      while (i * (i + 1) / 2 > e) --i;
      const int j = e - i * (i + 1) / 2;
      if (split_ok && !prepare) {
        Sc[i * ldM + j] = fma(rho, I.T[(size_t)b * tri + e], I.S0[(size_t)b * tri + e]);   // M(rho) = S0 + rho T
        continue;
      }
      const double *ai = At + i * ldT, *aj = At + j * ldT;
      if (prepare) {
        // S0 = P̄ + sigma I + rho_min A̅_f'A̅_f (free rows), T = A̅' diag(kappa) A̅ (kappa = 1 inequality, 1e3 equality)
        double s0 = gP[i * n + j] + (i == j ? S.sigma : 0.0), t0 = 0.0;
        for (int r = 0; r < m; ++r) {
          const int ct = row_class(lb[r], ub[r]);
          const int rt = r < mp ? r : r - mp;          // PAIRED: row r + mp is -row r, the products coincide
          const double aa = ai[rt] * aj[rt];
          if (ct < 0) s0 = fma(kRhoMin, aa, s0); else t0 = fma(ct == 1 ? kRhoEqOverIneq : 1.0, aa, t0);
        }
        I.S0[(size_t)b * tri + e] = s0; I.T[(size_t)b * tri + e] = t0;
        Sc[i * ldM + j] = fma(rho, t0, s0);
        continue;
      }
      double s0 = gP[i * n + j] + (i == j ? S.sigma : 0.0), s1 = 0.0;
      if (PAIRED) {
        for (int r = 0; r < mp; ++r) s0 = fma((rv[r] + rv[r + mp]) * ai[r], aj[r], s0);
      } else {
        for (int r = 0; r < mv; r += 2) {      // same summation order as admm_instance_kernel (pad column and rv pad are 0)
          const double2 rr = *reinterpret_cast<const double2 *>(rv + r), a = *reinterpret_cast<const double2 *>(ai + r), c2 = *reinterpret_cast<const double2 *>(aj + r);
          s0 = fma(rr.x * a.x, c2.x, s0);
          s1 = fma(rr.y * a.y, c2.y, s1);
        }
      }
      Sc[i * ldM + j] = s0 + s1;
    }
    __syncwarp();
    PFI(3)
    // left-looking Cholesky, lower triangle of Sc in place; zt[j] = 1 / L_jj
    int ok = 1;
    for (int j = 0; j < n; ++j) {
      if (lane >= j && lane < n) {
        const double *ri = Sc + lane * ldM, *rj = Sc + j * ldM;
        double s0 = ri[j], s1 = 0.0;
        int k = 0;
        for (; k + 2 <= j; k += 2) {
          const double2 a = *reinterpret_cast<const double2 *>(ri + k), c2 = *reinterpret_cast<const double2 *>(rj + k);
          s0 = fma(-a.x, c2.x, s0); s1 = fma(-a.y, c2.y, s1);
        }
        if (k < j) s0 = fma(-ri[k], rj[k], s0);
        xt[lane] = s0 + s1;    // column j before scaling (xt is free here)
      }
      __syncwarp();
      const double d = xt[j];
      if (!(d > 0.0)) { ok = 0; break; }
      const double sd = sqrt(d);
      if (lane >= j && lane < n) Sc[lane * ldM + j] = lane == j ? sd : xt[lane] / sd;
      if (lane == 0) zt[j] = 1.0 / sd;
      __syncwarp();
    }
    if (!ok) return false;
    PFI(4)
    // forward substitution L v = e_lane (v in mi[]), then backward L' x = v in place; L is read with warp-uniform addresses
#pragma unroll
    for (int i = 0; i < kRegN; ++i) {
      if (i < n) {
        double s0 = i == lane ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < i; k += 2) {
          const double2 l = *reinterpret_cast<const double2 *>(Sc + i * ldM + k);
          s0 = fma(-l.x, mi[k], s0); s1 = fma(-l.y, mi[k + 1], s1);
        }
        if (i & 1) s0 = fma(-Sc[i * ldM + i - 1], mi[i - 1], s0);
        mi[i] = (s0 + s1) * zt[i];
      } else {
        mi[i] = 0.0;
      }
    }
#pragma unroll
    for (int i = kRegN - 1; i >= 0; --i) {
      if (i < n) {
        double s0 = mi[i], s1 = 0.0;
#pragma unroll
        for (int k = i + 1; k < kRegN; ++k) {
          if (k < n) {
            if (k & 1) s1 = fma(-Sc[k * ldM + i], mi[k], s1); else s0 = fma(-Sc[k * ldM + i], mi[k], s0);
          }
        }
        mi[i] = (s0 + s1) * zt[i];
      }
    }
    if (lane >= n) {
#pragma unroll
      for (int k = 0; k < kRegN; ++k) mi[k] = 0.0;
    }
    __syncwarp();
    PFI(5)
    return true;
  };

  int status = SMPC_UNSOLVED, iter = 0;
  bool can_check = false;
  Info F = {};

  auto update_info = [&]() {
    for (int i = lane; i < n; i += 32) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s = fma(gP[k * n + i], x[k], s);   // P̄ x: P̄ is stored full symmetric, column sweep = coalesced
      sPx[i] = s;
    }
    {
      const double aty = AT_dot(y);
      if (lane < n) sAty[lane] = aty;
    }
    A_rows(x, w);
    double a_rp = 0, a_z = 0, a_Ax = 0, u_rp = 0, u_z = 0, u_Ax = 0;
    for (int r = lane; r < m; r += 32) {
      const double rp = w[r] - z[r], ei = Einv[r];
      a_rp = fmax(a_rp, fabs(rp)); a_z = fmax(a_z, fabs(z[r])); a_Ax = fmax(a_Ax, fabs(w[r]));
      u_rp = fmax(u_rp, fabs(ei * rp)); u_z = fmax(u_z, fabs(ei * z[r])); u_Ax = fmax(u_Ax, fabs(ei * w[r]));
    }
    double a_rd = 0, a_q = 0, a_Aty = 0, a_Px = 0, u_rd = 0, u_q = 0, u_Aty = 0, u_Px = 0, ob = 0;
    for (int i = lane; i < n; i += 32) {
      const double rd = (qb[i] + sPx[i]) + sAty[i], di = Dinv[i];
      a_rd = fmax(a_rd, fabs(rd)); a_q = fmax(a_q, fabs(qb[i])); a_Aty = fmax(a_Aty, fabs(sAty[i])); a_Px = fmax(a_Px, fabs(sPx[i]));
      u_rd = fmax(u_rd, fabs(di * rd)); u_q = fmax(u_q, fabs(di * qb[i])); u_Aty = fmax(u_Aty, fabs(di * sAty[i])); u_Px = fmax(u_Px, fabs(di * sPx[i]));
      ob += 0.5 * x[i] * sPx[i] + qb[i] * x[i];
    }
    F.s_rp = wmax(a_rp); F.s_z = wmax(a_z); F.s_Ax = wmax(a_Ax);
    F.s_rd = wmax(a_rd); F.s_q = wmax(a_q); F.s_Aty = wmax(a_Aty); F.s_Px = wmax(a_Px);
    if (unscale) {
      F.pri_res = wmax(u_rp); F.nEz = wmax(u_z); F.nEAx = wmax(u_Ax);
      F.dua_res = cinv * wmax(u_rd); F.nDq = wmax(u_q); F.nDAty = wmax(u_Aty); F.nDPx = wmax(u_Px);
      F.obj = cinv * wsum(ob);
    } else {
      F.pri_res = F.s_rp; F.nEz = F.s_z; F.nEAx = F.s_Ax;
      F.dua_res = F.s_rd; F.nDq = F.s_q; F.nDAty = F.s_Aty; F.nDPx = F.s_Px;
      F.obj = wsum(ob);
    }
    if (m == 0) F.pri_res = 0.0;
  };

  auto primal_infeasible = [&](double eps) -> bool {
    double nd = 0.0;
    for (int r = lane; r < m; r += 32) {
      double d = dy[r];
      const bool uinf = ub[r] > kInfty * kMinScaling, linf = lb[r] < -kInfty * kMinScaling;
      if (uinf) d = linf ? 0.0 : fmin(d, 0.0); else if (linf) d = fmax(d, 0.0);
      dy[r] = d;
      nd = fmax(nd, fabs(unscale ? Ev[r] * d : d));
    }
    nd = wmax(nd);
    __syncwarp();
    if (!(nd > eps)) return false;
    double lhs = 0.0;
    for (int r = lane; r < m; r += 32) {
      const double dp = fmax(dy[r], 0.0), dm = fmin(dy[r], 0.0);
      if (dp != 0.0) lhs += ub[r] * dp;
      if (dm != 0.0) lhs += lb[r] * dm;
    }
    lhs = wsum(lhs);
    if (!(lhs < -eps * nd)) return false;
    double na = 0.0;
    {
      const double v = AT_dot(dy);
      if (lane < n) na = fabs(unscale ? Dinv[lane] * v : v);
    }
    return wmax(na) < eps * nd;
  };

  auto dual_infeasible = [&](double eps) -> bool {
    double nd = 0.0, qd = 0.0;
    for (int i = lane; i < n; i += 32) {
      nd = fmax(nd, fabs(unscale ? Dv[i] * dx[i] : dx[i]));
      qd += qb[i] * dx[i];
    }
    nd = wmax(nd); qd = wsum(qd);
    const double cs = unscale ? c : 1.0;
    if (!(nd > eps)) return false;
    if (!(qd < -cs * eps * nd)) return false;
    double np = 0.0;
    for (int i = lane; i < n; i += 32) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s = fma(gP[k * n + i], dx[k], s);
      np = fmax(np, fabs(unscale ? Dinv[i] * s : s));
    }
    if (!(wmax(np) < cs * eps * nd)) return false;
    int bad = 0;
    A_rows(dx, w);
    for (int r = lane; r < m; r += 32) {
      double v = w[r];
      if (unscale) v *= Einv[r];
      if (((ub[r] < kInfty * kMinScaling) && (v > eps * nd)) || ((lb[r] > -kInfty * kMinScaling) && (v < -eps * nd))) bad = 1;
    }
    return !__any_sync(kFull, bad);
  };

  auto check_termination = [&](bool approx) -> bool {
    double ea = S.eps_abs, er = S.eps_rel, epi = S.eps_prim_inf, edi = S.eps_dual_inf;
    if (approx) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (m == 0) prim_ok = true;
    else if (F.pri_res < ea + er * fmax(F.nEz, F.nEAx)) prim_ok = true;
    else prim_inf = primal_infeasible(epi);
    if (F.dua_res < ea + er * (unscale ? cinv : 1.0) * fmax(fmax(F.nDq, F.nDAty), F.nDPx)) dual_ok = true;
    else dual_inf = dual_infeasible(edi);
    if (prim_ok && dual_ok) { status = approx ? SMPC_SOLVED_INACCURATE : SMPC_SOLVED; return true; }
    if (prim_inf) { status = approx ? SMPC_PRIMAL_INFEASIBLE_INACCURATE : SMPC_PRIMAL_INFEASIBLE; F.obj = kInfty; return true; }
    if (dual_inf) { status = approx ? SMPC_DUAL_INFEASIBLE_INACCURATE : SMPC_DUAL_INFEASIBLE; F.obj = -kInfty; return true; }
    return false;
  };

  PFI(0)
  bool factor_ok = true, need_factor = true;   // (re)factor at the top of the iteration that first needs it: one call site
  for (iter = 1; iter <= S.max_iter && !bad_bounds && factor_ok; ++iter) {
    if (need_factor) {
      set_rho_vec();
      if (split_ok && !prepare && rho == I.rho_prepared) {
        // the factorisation of osqp_setup: rows of M(rho0)^-1 prepared at create time
#pragma unroll
        for (int k = 0; k < kRegN; ++k) mi[k] = lane < n ? I.Minv0[((size_t)b * kRegN + k) * n + lane] : 0.0;   // [B][32][n]: coalesced over the lanes
        factor_ok = true;
      } else {
        factor_ok = refactor();
      }
      need_factor = false;
      if (prepare) {
        if (factor_ok && lane < n) {
#pragma unroll
          for (int k = 0; k < kRegN; ++k) I.Minv0[((size_t)b * kRegN + k) * n + lane] = mi[k];
        }
        return;
      }
      if (!factor_ok) break;
      for (int r = lane; r < m; r += 32) zt[r] = 1.0 / rv[r];   // 1 / rho_vec for the z update (zt was refactor scratch)
      __syncwarp();
    }
    for (int r = lane; r < m; r += 32) w[r] = rv[r] * z[r] - y[r];
    __syncwarp();
    {
      const double aw = AT_dot(w);
      if (lane < n) sAty[lane] = (S.sigma * x[lane] - qb[lane]) + aw;                                     // rhs
    }
    __syncwarp();
    {
      const double xv = dot_reg32(mi, sAty);                                                              // x̃ = M^-1 rhs
      if (lane < n) xt[lane] = xv;
    }
    __syncwarp();
    A_rows(xt, w);                                                                                        // z̃ = A̅ x̃ (w is free after the rhs)
    for (int r = lane; r < m; r += 32) {
      const double ztl = w[r];
      const double rr = rv[r], rinv = zt[r];
      const double zr = alpha * ztl + (1.0 - alpha) * z[r];
      const double zn = fmin(fmax(zr + rinv * y[r], lb[r]), ub[r]);
      const double d = rr * (zr - zn);
      z[r] = zn; y[r] += d; dy[r] = d;
    }
    for (int i = lane; i < n; i += 32) {
      const double xn = alpha * xt[i] + (1.0 - alpha) * x[i];
      dx[i] = xn - x[i];
      x[i] = xn;
    }
    __syncwarp();
    PFI(1)
    can_check = S.check_every && (iter % S.check_every == 0);
    if (can_check) {
      update_info();
      if (check_termination(false)) break;
    }
    if (S.adaptive_rho && S.rho_interval && (iter % S.rho_interval == 0)) {
      if (!can_check) update_info();
      const double pr = F.s_rp / (fmax(F.s_z, F.s_Ax) + kDivTol);
      const double dr = F.s_rd / (fmax(fmax(F.s_q, F.s_Aty), F.s_Px) + kDivTol);
      const double rn = fmin(fmax(rho * sqrt(pr / (dr + kDivTol)), kRhoMin), kRhoMax);
      if (rn > rho * S.rho_tol || rn < rho / S.rho_tol) {
        rho = rn; ++rho_updates;
        need_factor = true;
      }
    }
    __syncwarp();
    PFI(6)
  }
  if (iter > S.max_iter) iter = S.max_iter;
  if (bad_bounds || !factor_ok) iter = 0;
  else {
    if (!can_check) { update_info(); check_termination(false); }
    if (status == SMPC_UNSOLVED) { if (!check_termination(true)) status = SMPC_MAX_ITER_REACHED; }
  }

  const bool has_sol = !bad_bounds && factor_ok && !(status == SMPC_PRIMAL_INFEASIBLE || status == SMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                                                     status == SMPC_DUAL_INFEASIBLE || status == SMPC_DUAL_INFEASIBLE_INACCURATE);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    if (Bt.x_out) Bt.x_out[(size_t)b * n + i] = has_sol ? Dv[i] * x[i] : qnan;
    Bt.xi[(size_t)b * n + i] = has_sol ? x[i] : 0.0;
  }
  for (int r = lane; r < m; r += 32) {
    if (Bt.y_out) Bt.y_out[(size_t)b * m + r] = has_sol ? cinv * (Ev[r] * y[r]) : qnan;
    Bt.z[(size_t)b * m + r] = has_sol ? z[r] : 0.0;
    Bt.y[(size_t)b * m + r] = has_sol ? y[r] : 0.0;
  }
#ifdef SMPC_PROFILE
  PFI(7)
  if (b == 0 && lane == 0) printf("instance profile (cycles): load %lld, iterations %lld, pre-assembly %lld, M assembly %lld, cholesky %lld, tri solves %lld, checks %lld, tail %lld; iters %d rho updates %d\n", pf[0], pf[1], pf[2], pf[3], pf[4], pf[5], pf[6], pf[7], iter, rho_updates);
#endif
  if (lane == 0) {
    Bt.rho[b] = rho;
    Bt.status[b] = status; Bt.iter[b] = iter; Bt.rho_updates[b] = rho_updates;
    Bt.obj[b] = F.obj; Bt.pri_res[b] = F.pri_res; Bt.dua_res[b] = F.dua_res;
  }
}

// osqp_warm_start in the per-instance regime: x̄ = D^-1 x, z = A̅ x̄, ȳ = c E^-1 y
__global__ void warm_start_instance_kernel(InstanceDataDev I, const double *x, const double *y, double *xs, double *z, double *ys) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= I.B) return;
  const int n = I.n, m = I.m;
  const double c = I.c[b];
  for (int i = lane; i < n; i += 32) xs[(size_t)b * n + i] = x[(size_t)b * n + i] / I.D[(size_t)b * n + i];
  __syncwarp();
  for (int r = lane; r < m; r += 32) {
    double s = 0.0;
    for (int k = 0; k < n; ++k) s = fma(I.A[((size_t)b * m + r) * n + k], xs[(size_t)b * n + k], s);
    z[(size_t)b * m + r] = s;
    ys[(size_t)b * m + r] = c * (y[(size_t)b * m + r] / I.E[(size_t)b * m + r]);
  }
}

static int pick_wpc(size_t per_warp_bytes) {
  int wpc = 8;
  while (wpc > 1 && wpc * per_warp_bytes > 224 * 1024) --wpc;
  return wpc;
}

bool instance_kernel_supports(int n, int m) {
  return n >= 1 && m >= 0 && instance_warp_doubles(n, m) * sizeof(double) <= 227 * 1024 &&
         ((size_t)(n + m) * odd_ld(n) + 2 * n + 2 * m) * sizeof(double) <= 227 * 1024;
}

cudaError_t launch_ruiz_instance(const InstanceDataDev &I, int iters, cudaStream_t stream) {
  const size_t per = ((size_t)(I.n + I.m) * odd_ld(I.n) + 2 * I.n + 2 * I.m) * sizeof(double);
  const int wpc = pick_wpc(per);
  const size_t smem = wpc * per;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ruiz_instance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  ruiz_instance_kernel<<<(I.B + wpc - 1) / wpc, wpc * 32, smem, stream>>>(I, iters, wpc);
  return cudaGetLastError();
}

// flag[0] &= "rows r and r + m/2 of every (scaled) A̅_i are exact negatives of each other"
__global__ void instance_pairs_kernel(InstanceDataDev I, int *flag) {
  const int mp = I.m / 2;
  const size_t total = (size_t)I.B * mp * I.n;
  int ok = 1;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t bq = e / ((size_t)mp * I.n), rem = e % ((size_t)mp * I.n);
    const double *A = I.A + bq * I.m * I.n;
    ok &= A[rem + (size_t)mp * I.n] == -A[rem];
  }
  if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) atomicAnd(flag, 0);
}
cudaError_t launch_instance_pairs(const InstanceDataDev &I, int *flag, cudaStream_t stream) {
  instance_pairs_kernel<<<148 * 4, 256, 0, stream>>>(I, flag);
  return cudaGetLastError();
}

bool instance_reg_supports(int n, int m) { return n >= 1 && n <= kRegN && m >= 1 && m <= kRegM; }

cudaError_t launch_admm_instance(const InstanceDataDev &I, const BatchDev &Bt, const SettingsDev &S, cudaStream_t stream, int prepare) {
  if (prepare && !(instance_reg_supports(I.n, I.m) && I.S0)) return cudaSuccess;   // nothing to prepare for the generic kernel
  if (instance_reg_supports(I.n, I.m)) {   // register-operator variant
    const bool paired = I.paired != 0;
    const size_t per = ((instance_reg_warp_doubles(I.n, I.m, paired) + 1) & ~(size_t)1) * sizeof(double);
    const int wpc = pick_wpc(per);
    const size_t smem = wpc * per;
    // (the attribute is per device: set it on every launch rather than once per process)
    cudaError_t e = paired ? cudaFuncSetAttribute(admm_instance_reg_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                           : cudaFuncSetAttribute(admm_instance_reg_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (paired) admm_instance_reg_kernel<true><<<(Bt.B + wpc - 1) / wpc, wpc * 32, smem, stream>>>(I, Bt, S, wpc, prepare);
    else admm_instance_reg_kernel<false><<<(Bt.B + wpc - 1) / wpc, wpc * 32, smem, stream>>>(I, Bt, S, wpc, prepare);
    return cudaGetLastError();
  }
  const size_t per = instance_warp_doubles(I.n, I.m) * sizeof(double);
  const int wpc = pick_wpc(per);
  const size_t smem = wpc * per;
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(admm_instance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  admm_instance_kernel<<<(Bt.B + wpc - 1) / wpc, wpc * 32, smem, stream>>>(I, Bt, S, wpc);
  return cudaGetLastError();
}

cudaError_t launch_warm_start_instance(const InstanceDataDev &I, const double *x, const double *y, double *xs, double *z,
                                       double *ys, cudaStream_t stream) {
  const int wpc = 4;
  warm_start_instance_kernel<<<(I.B + wpc - 1) / wpc, wpc * 32, 0, stream>>>(I, x, y, xs, z, ys);
  return cudaGetLastError();
}

}  // namespace smpc
