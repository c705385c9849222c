// admm_shared_generic.cu -- generic-size resident ADMM kernel of the shared-factor regime.
//
// One warp owns one QP from its first to its last ADMM iteration: q, l, u are read from HBM
// once, the iterates (xi, z, y) live in the warp's shared-memory slice for the whole solve and
// x, y, status are written once.  Nothing per-iteration touches HBM; the shared operators
// (sigma*G, W, W', ...) are read through L1/L2 where they stay resident (<= a few hundred KB).
// Per-problem convergence masking is the warp leaving its loop; rho adaptation is per problem
// and needs no refactorisation (plan.hpp).  The iteration restates OSQP 0.6.x osqp_solve
// (SURVEY.md 3.4), which the reference runs through solver.solve()
// (src/ModelPredictiveControlAPI.cpp:102), in the coordinates x̄ = V xi:
//     w   = rho_vec .* z - y
//     t   = (sigma*G xi - q̂ + W' w) ./ (1 + rho*lambda)          [ = V^-1 x̃ ]
//     z̃   = W t
//     xi  = alpha t + (1-alpha) xi
//     z   = clip(alpha z̃ + (1-alpha) z + y ./ rho_vec, l̄, ū);  y += rho_vec .* (alpha z̃ + (1-alpha) z_prev - z)
// Lanes own rows (lane-strided), vectors are broadcast from shared memory.
#include "device_types.cuh"
#include "kernels.cuh"

namespace smpc {

namespace {

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_any(int p) { return __any_sync(0xffffffffu, p); }

// sum_k MT[k*ld + i] * vec[k], k in [0,K)
__device__ __forceinline__ double col_dot(const double *__restrict__ MT, int ld, int K, int i,
                                          const double *vec) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int k = 0;
  for (; k + 4 <= K; k += 4) {
    a0 = fma(__ldg(MT + (size_t)(k + 0) * ld + i), vec[k + 0], a0);
    a1 = fma(__ldg(MT + (size_t)(k + 1) * ld + i), vec[k + 1], a1);
    a2 = fma(__ldg(MT + (size_t)(k + 2) * ld + i), vec[k + 2], a2);
    a3 = fma(__ldg(MT + (size_t)(k + 3) * ld + i), vec[k + 3], a3);
  }
  for (; k < K; ++k) a0 = fma(__ldg(MT + (size_t)k * ld + i), vec[k], a0);
  return (a0 + a1) + (a2 + a3);
}

__device__ __forceinline__ double rho_of_row(int ct, double rho) {
  return ct == 0 ? rho : (ct == 1 ? kRhoEqOverIneq * rho : kRhoMin);
}

struct Info {  // what OSQP's update_info leaves behind (uniform across the warp)
  double pri_res, dua_res, nEz, nEAx, nDq, nDAty, nDPx;   // unscaled (or scaled if scaled_termination)
  double s_rp, s_rd, s_z, s_Ax, s_q, s_Aty, s_Px;          // scaled norms for the rho estimate
  double obj;
};

}  // namespace

__global__ void __launch_bounds__(256) admm_shared_generic_kernel(SharedPlanDev P, BatchDev Bt,
                                                                  SettingsDev S, int warps_per_cta) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * warps_per_cta + warp;
  if (b >= Bt.B) return;
  const int n = P.n, m = P.m;
  double *base = smem + (size_t)warp * 7 * (n + m);
  double *xi = base, *dxi = xi + n, *qh = dxi + n, *qb = qh + n, *t = qb + n, *sPx = t + n, *sAty = sPx + n;
  double *z = sAty + n, *y = z + m, *lb = y + m, *ub = lb + m, *w = ub + m, *zt = w + m, *dy = zt + m;
  const double alpha = S.alpha, c = P.c, cinv = P.cinv;
  const bool unscale = !S.scaled_termination;

  // ---- load the instance: q̄ = c D q (osqp_update_lin_cost), l̄ = E l, ū = E u (osqp_update_bounds)
  for (int i = lane; i < n; i += 32) {
    qb[i] = Bt.q ? c * (P.D[i] * Bt.q[(size_t)b * n + i]) : 0.0;
    xi[i] = (S.warm_start && !Bt.fresh) ? Bt.xi[(size_t)b * n + i] : 0.0;
    dxi[i] = 0.0;
  }
  for (int r = lane; r < m; r += 32) {
    double lo = Bt.l ? Bt.l[(size_t)b * m + r] : P.l0[r];
    double hi = Bt.u ? Bt.u[(size_t)b * m + r] : P.u0[r];
    lb[r] = P.E[r] * lo; ub[r] = P.E[r] * hi;
    z[r] = (S.warm_start && !Bt.fresh) ? Bt.z[(size_t)b * m + r] : 0.0;
    y[r] = (S.warm_start && !Bt.fresh) ? Bt.y[(size_t)b * m + r] : 0.0;
    dy[r] = 0.0;
  }
  double rho = Bt.fresh ? fmin(fmax(S.rho0, kRhoMin), kRhoMax) : Bt.rho[b];
  int rho_updates = 0;
  __syncwarp();
  // osqp_update_bounds refuses l > u; such an instance is left UNSOLVED (iter 0, NaN solution).
  // An instance whose bounds change a row's class (equality / inequality / free) against the SETUP bounds: OSQP would re-classify
  // the row and refactor for that instance alone, which the shared factor cannot follow.  It is solved with the plan's rho_vec
  // entries instead -- rho_vec is any positive diagonal for ADMM, the fixed point and every termination test are the same, only
  // the iterate path (and so the iteration count) can differ from osqp-eigen's for that instance.
  int bad_rows = 0;
  for (int r = lane; r < m; r += 32) bad_rows |= (lb[r] > ub[r]);
  const bool bad_bounds = warp_any(bad_rows);
  for (int i = lane; i < n; i += 32) qh[i] = col_dot(P.V, n, n, i, qb);   // q̂ = V' q̄
  __syncwarp();

  int status = SMPC_UNSOLVED, iter = 0;
  bool can_check = false;
  Info I = {};

  // OSQP update_info: residuals, tolerances' norms, objective.  Leaves x̄ in t, P̄x in sPx, A̅'y in sAty, A̅x in w.
  auto update_info = [&]() {
    for (int i = lane; i < n; i += 32) {
      t[i] = col_dot(P.VT, n, n, i, xi);
      sPx[i] = col_dot(P.PVT, n, n, i, xi);
      sAty[i] = col_dot(P.Abar, n, m, i, y);
    }
    for (int r = lane; r < m; r += 32) w[r] = col_dot(P.WT, m, n, r, xi);
    __syncwarp();
    double a_rp = 0, a_z = 0, a_Ax = 0, u_rp = 0, u_z = 0, u_Ax = 0;
    for (int r = lane; r < m; r += 32) {
      double rp = w[r] - z[r], ei = P.Einv[r];
      a_rp = fmax(a_rp, fabs(rp)); a_z = fmax(a_z, fabs(z[r])); a_Ax = fmax(a_Ax, fabs(w[r]));
      u_rp = fmax(u_rp, fabs(ei * rp)); u_z = fmax(u_z, fabs(ei * z[r])); u_Ax = fmax(u_Ax, fabs(ei * w[r]));
    }
    double a_rd = 0, a_q = 0, a_Aty = 0, a_Px = 0, u_rd = 0, u_q = 0, u_Aty = 0, u_Px = 0, ob = 0;
    for (int i = lane; i < n; i += 32) {
      double rd = (qb[i] + sPx[i]) + sAty[i], di = P.Dinv[i];
      a_rd = fmax(a_rd, fabs(rd)); a_q = fmax(a_q, fabs(qb[i])); a_Aty = fmax(a_Aty, fabs(sAty[i])); a_Px = fmax(a_Px, fabs(sPx[i]));
      u_rd = fmax(u_rd, fabs(di * rd)); u_q = fmax(u_q, fabs(di * qb[i])); u_Aty = fmax(u_Aty, fabs(di * sAty[i])); u_Px = fmax(u_Px, fabs(di * sPx[i]));
      ob += 0.5 * t[i] * sPx[i] + qb[i] * t[i];
    }
    I.s_rp = warp_max(a_rp); I.s_z = warp_max(a_z); I.s_Ax = warp_max(a_Ax);
    I.s_rd = warp_max(a_rd); I.s_q = warp_max(a_q); I.s_Aty = warp_max(a_Aty); I.s_Px = warp_max(a_Px);
    if (unscale) {
      I.pri_res = warp_max(u_rp); I.nEz = warp_max(u_z); I.nEAx = warp_max(u_Ax);
      I.dua_res = cinv * warp_max(u_rd); I.nDq = warp_max(u_q); I.nDAty = warp_max(u_Aty); I.nDPx = warp_max(u_Px);
      I.obj = cinv * warp_sum(ob);
    } else {
      I.pri_res = I.s_rp; I.nEz = I.s_z; I.nEAx = I.s_Ax;
      I.dua_res = I.s_rd; I.nDq = I.s_q; I.nDAty = I.s_Aty; I.nDPx = I.s_Px;
      I.obj = warp_sum(ob);
    }
    if (m == 0) I.pri_res = 0.0;
  };

  // OSQP is_primal_infeasible (uses the last delta_y; projection onto the recession-cone polar first)
  auto primal_infeasible = [&](double eps) -> bool {
    double nd = 0.0;
    for (int r = lane; r < m; r += 32) {
      double d = dy[r];
      bool uinf = ub[r] > kInfty * kMinScaling, linf = lb[r] < -kInfty * kMinScaling;
      if (uinf) d = linf ? 0.0 : fmin(d, 0.0); else if (linf) d = fmax(d, 0.0);
      dy[r] = d;
      nd = fmax(nd, fabs(unscale ? P.E[r] * d : d));
    }
    nd = warp_max(nd);
    __syncwarp();
    if (!(nd > eps)) return false;
    double lhs = 0.0;
    for (int r = lane; r < m; r += 32) {
      double dp = fmax(dy[r], 0.0), dm = fmin(dy[r], 0.0);
      if (dp != 0.0) lhs += ub[r] * dp;
      if (dm != 0.0) lhs += lb[r] * dm;
    }
    lhs = warp_sum(lhs);
    if (!(lhs < -eps * nd)) return false;
    double na = 0.0;
    for (int i = lane; i < n; i += 32) {
      double v = col_dot(P.Abar, n, m, i, dy);
      na = fmax(na, fabs(unscale ? P.Dinv[i] * v : v));
    }
    return warp_max(na) < eps * nd;
  };

  // OSQP is_dual_infeasible (uses the last delta_x = V delta_xi)
  auto dual_infeasible = [&](double eps) -> bool {
    double nd = 0.0, qd = 0.0;
    for (int i = lane; i < n; i += 32) {
      double dx = col_dot(P.VT, n, n, i, dxi);
      nd = fmax(nd, fabs(unscale ? P.D[i] * dx : dx));
      qd += qb[i] * dx;
    }
    nd = warp_max(nd); qd = warp_sum(qd);
    double cs = unscale ? c : 1.0;
    if (!(nd > eps)) return false;
    if (!(qd < -cs * eps * nd)) return false;
    double np = 0.0;
    for (int i = lane; i < n; i += 32) {
      double v = col_dot(P.PVT, n, n, i, dxi);
      np = fmax(np, fabs(unscale ? P.Dinv[i] * v : v));
    }
    if (!(warp_max(np) < cs * eps * nd)) return false;
    int bad = 0;
    for (int r = lane; r < m; r += 32) {
      double v = col_dot(P.WT, m, n, r, dxi);
      if (unscale) v *= P.Einv[r];
      if (((ub[r] < kInfty * kMinScaling) && (v > eps * nd)) || ((lb[r] > -kInfty * kMinScaling) && (v < -eps * nd))) bad = 1;
    }
    return !warp_any(bad);
  };

  // OSQP check_termination; returns true when the status was set
  auto check_termination = [&](bool approx) -> bool {
    double ea = S.eps_abs, er = S.eps_rel, epi = S.eps_prim_inf, edi = S.eps_dual_inf;
    if (approx) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (m == 0) prim_ok = true;
    else {
      double ep = ea + er * fmax(I.nEz, I.nEAx);
      if (I.pri_res < ep) prim_ok = true; else prim_inf = primal_infeasible(epi);
    }
    double ed = ea + er * (unscale ? cinv : 1.0) * fmax(fmax(I.nDq, I.nDAty), I.nDPx);
    if (I.dua_res < ed) dual_ok = true; else dual_inf = dual_infeasible(edi);
    if (prim_ok && dual_ok) { status = approx ? SMPC_SOLVED_INACCURATE : SMPC_SOLVED; return true; }
    if (prim_inf) { status = approx ? SMPC_PRIMAL_INFEASIBLE_INACCURATE : SMPC_PRIMAL_INFEASIBLE; I.obj = kInfty; return true; }
    if (dual_inf) { status = approx ? SMPC_DUAL_INFEASIBLE_INACCURATE : SMPC_DUAL_INFEASIBLE; I.obj = -kInfty; return true; }
    return false;
  };

  for (iter = 1; iter <= S.max_iter && !bad_bounds; ++iter) {
    // w = rho_vec z - y
    for (int r = lane; r < m; r += 32) w[r] = rho_of_row(P.ctype[r], rho) * z[r] - y[r];
    __syncwarp();
    // t = (sigma G xi + W' w - q̂) / (1 + rho lambda)
    for (int i = lane; i < n; i += 32) {
      double acc = col_dot(P.SG, n, n, i, xi) + col_dot(P.W, n, m, i, w) - qh[i];
      t[i] = acc / (1.0 + rho * P.lam[i]);
    }
    __syncwarp();
    // z̃ = W t ; x / z / y updates
    for (int r = lane; r < m; r += 32) {
      double ztl = col_dot(P.WT, m, n, r, t);
      double rv = rho_of_row(P.ctype[r], rho), rinv = 1.0 / rv;
      double zr = alpha * ztl + (1.0 - alpha) * z[r];
      double zn = fmin(fmax(zr + rinv * y[r], lb[r]), ub[r]);
      double d = rv * (zr - zn);
      z[r] = zn; y[r] += d; dy[r] = d;
    }
    for (int i = lane; i < n; i += 32) {
      double xn = alpha * t[i] + (1.0 - alpha) * xi[i];
      dxi[i] = xn - xi[i];
      xi[i] = xn;
    }
    __syncwarp();
    can_check = S.check_every && (iter % S.check_every == 0);
    if (can_check) {
      update_info();
      if (check_termination(false)) break;
    }
    if (S.adaptive_rho && S.rho_interval && (iter % S.rho_interval == 0)) {
      if (!can_check) update_info();
      // OSQP compute_rho_estimate / adapt_rho (scaled norms)
      double pr = I.s_rp / (fmax(I.s_z, I.s_Ax) + kDivTol);
      double dr = I.s_rd / (fmax(fmax(I.s_q, I.s_Aty), I.s_Px) + kDivTol);
      double rn = fmin(fmax(rho * sqrt(pr / (dr + kDivTol)), kRhoMin), kRhoMax);
      if (rn > rho * S.rho_tol || rn < rho / S.rho_tol) { rho = rn; ++rho_updates; }
    }
    __syncwarp();
  }
  if (iter > S.max_iter) iter = S.max_iter;
  if (bad_bounds) iter = 0;
  else {
    if (!can_check) { update_info(); check_termination(false); }
    if (status == SMPC_UNSOLVED) { if (!check_termination(true)) status = SMPC_MAX_ITER_REACHED; }
  }

  // ---- store_solution: x = D x̄ (x̄ = V xi is in t), y = E ȳ / c ; infeasible -> NaN and cold start
  const bool has_sol = !bad_bounds && !(status == SMPC_PRIMAL_INFEASIBLE || status == SMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                         status == SMPC_DUAL_INFEASIBLE || status == SMPC_DUAL_INFEASIBLE_INACCURATE);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    if (Bt.x_out) Bt.x_out[(size_t)b * n + i] = has_sol ? P.D[i] * t[i] : qnan;
    Bt.xi[(size_t)b * n + i] = has_sol ? xi[i] : 0.0;
  }
  for (int r = lane; r < m; r += 32) {
    if (Bt.y_out) Bt.y_out[(size_t)b * m + r] = has_sol ? cinv * (P.E[r] * y[r]) : qnan;
    Bt.z[(size_t)b * m + r] = has_sol ? z[r] : 0.0;
    Bt.y[(size_t)b * m + r] = has_sol ? y[r] : 0.0;
  }
  if (lane == 0) {
    Bt.rho[b] = rho;
    Bt.status[b] = status; Bt.iter[b] = iter; Bt.rho_updates[b] = rho_updates;
    Bt.obj[b] = I.obj; Bt.pri_res[b] = I.pri_res; Bt.dua_res[b] = I.dua_res;
  }
}

// ---- small helper kernels -----------------------------------------------------------------
__global__ void fill_kernel(double *p, double v, size_t count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) p[i] = v;
}

// osqp_warm_start: x̄ = D^-1 x, xi = Vinv x̄, z = A̅ x̄ = W xi, ȳ = c E^-1 y.  One warp per instance.
__global__ void warm_start_kernel(SharedPlanDev P, int B, const double *x, const double *y,
                                  double *xi, double *z, double *ys, int xspace) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  const int b = blockIdx.x * wpc + warp;
  if (b >= B) return;
  const int n = P.n, m = P.m;
  double *xb = smem + (size_t)warp * 2 * n, *xv = xb + n;
  for (int i = lane; i < n; i += 32) xb[i] = P.Dinv[i] * x[(size_t)b * n + i];
  __syncwarp();
  for (int i = lane; i < n; i += 32) { double s = 0; for (int k = 0; k < n; ++k) s = fma(P.VinvT[(size_t)k * n + i], xb[k], s); xv[i] = s; }
  __syncwarp();
  for (int i = lane; i < n; i += 32) xi[(size_t)b * n + i] = xspace ? xb[i] : xv[i];   // x-space kernels keep x̄ itself
  for (int r = lane; r < m; r += 32) {
    double s = 0; for (int k = 0; k < n; ++k) s = fma(P.WT[(size_t)k * m + r], xv[k], s);
    z[(size_t)b * m + r] = s;
    ys[(size_t)b * m + r] = P.c * (P.Einv[r] * y[(size_t)b * m + r]);
  }
}

size_t generic_smem_bytes(int n, int m, int warps_per_cta) { return (size_t)warps_per_cta * 7 * (n + m) * sizeof(double); }

cudaError_t launch_admm_shared_generic(const SharedPlanDev &P, const BatchDev &Bt, const SettingsDev &S,
                                       cudaStream_t stream) {
  int wpc = 8;
  while (wpc > 1 && generic_smem_bytes(P.n, P.m, wpc) > 200 * 1024) --wpc;
  size_t smem = generic_smem_bytes(P.n, P.m, wpc);
  if (smem > 227 * 1024) return cudaErrorInvalidValue;
  if (smem > 48 * 1024) {   // the attribute is per device: set it on every launch (cheap), as the other launchers do
    cudaError_t e = cudaFuncSetAttribute(admm_shared_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  int grid = (Bt.B + wpc - 1) / wpc;
  admm_shared_generic_kernel<<<grid, wpc * 32, smem, stream>>>(P, Bt, S, wpc);
  return cudaGetLastError();
}

cudaError_t launch_fill(double *p, double v, size_t count, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  fill_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(p, v, count);
  return cudaGetLastError();
}

cudaError_t launch_warm_start(const SharedPlanDev &P, int B, const double *x, const double *y, double *xi,
                              double *z, double *ys, cudaStream_t stream, bool xspace) {
  int wpc = 4;
  size_t smem = (size_t)wpc * 2 * P.n * sizeof(double);
  warm_start_kernel<<<(B + wpc - 1) / wpc, wpc * 32, smem, stream>>>(P, B, x, y, xi, z, ys, xspace ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace smpc
