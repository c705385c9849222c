// admm_shared_small.cu -- register-resident ADMM kernel for small QPs (n <= 16, m <= 32), the shape of
// the reference's own problem (n = 15, m = 30: mpcWindow = 15, include/ModelPredictiveControlAPI.h:26).
//
// A persistent grid of warps pulls QPs from a device-side queue.  One warp owns one QP from its first
// ADMM iteration to its last; the iterates AND the shared operators live in registers:
//   lane (h, i) = (lane >> 4, lane & 15) keeps  m1[j] = [sigma*G | W'](i, 24h + j), j < 24   (t-phase, K split in 2)
//   lane r      = lane                    keeps  wr[i] = W(r, i), i < 16                      (z-phase)
//   and its rows of xi, q̂, 1/(1+rho*lambda), z, y, l̄, ū.
// Per iteration (OSQP 0.6.x osqp_solve in plan coordinates, see admm_shared_generic.cu / plan.hpp):
//   24 DFMA + 1 shuffle-add for t, 16 DFMA for z̃, ~13 FP64 element-wise ops, two 512-byte shared-memory
//   broadcasts ([xi; w] and t).  HBM is touched once per solve (q, u in; x, y, status out), so the kernel is
//   bound by the FP64 pipe, not by HBM.
// Termination checks / rho adaptation (every 25 iterations) read their operators from shared memory.
#include "device_types.cuh"
#include "kernels.cuh"

namespace smpc {

namespace {

constexpr int NP = 16, MP = 32, KH = (NP + MP) / 2;   // padded sizes; 24 concatenated entries per half-warp

__device__ __forceinline__ double wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double rho_row(int ct, double rho) {
  return ct == 0 ? rho : (ct == 1 ? kRhoEqOverIneq * rho : kRhoMin);
}

struct Info {
  double pri_res, dua_res, nEz, nEAx, nDq, nDAty, nDPx;
  double s_rp, s_rd, s_z, s_Ax, s_q, s_Aty, s_Px;
  double obj;
};

}  // namespace

// CTA-shared operator block (doubles): VT[16][16] PVT[16][16] AbP[32][16] V[16][16]
constexpr int kCtaMatDoubles = NP * NP * 3 + MP * NP;
// per-warp block (doubles): cbuf[48] tbuf[16] sbuf[32]
constexpr int kWarpDoubles = NP + MP + NP + MP;

__global__ void __launch_bounds__(128, 3)
admm_shared_small_kernel(SmallPackDev K, SharedPlanDev P, BatchDev Bt, SettingsDev S, int *queue) {
  extern __shared__ __align__(16) double smem[];
  double *sVT = smem, *sPVT = sVT + NP * NP, *sAb = sPVT + NP * NP, *sV = sAb + MP * NP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *cbuf = smem + kCtaMatDoubles + warp * kWarpDoubles, *tbuf = cbuf + NP + MP, *sbuf = tbuf + NP;
  const int h = lane >> 4, i = lane & 15, r = lane;
  const int n = P.n, m = P.m;

  for (int e = threadIdx.x; e < NP * NP; e += blockDim.x) { sVT[e] = K.VT[e]; sPVT[e] = K.PVT[e]; sV[e] = K.V[e]; }
  for (int e = threadIdx.x; e < MP * NP; e += blockDim.x) sAb[e] = K.Ab[e];

  // shared operators -> registers (coalesced: the packs are stored k-major)
  double m1[KH], wr[NP];
#pragma unroll
  for (int j = 0; j < KH; ++j) m1[j] = K.M1T[(KH * h + j) * NP + i];
#pragma unroll
  for (int k = 0; k < NP; ++k) wr[k] = K.WT[k * MP + r];
  const double lam_i = K.lam[i], D_i = K.D[i], Dinv_i = K.Dinv[i];
  const double E_r = K.E[r], Einv_r = K.Einv[r];
  const int ct_r = K.ctype[r];
  const double alpha = S.alpha, c = P.c, cinv = P.cinv;
  const bool unscale = !S.scaled_termination;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  __syncthreads();

  for (;;) {
    int b = 0;
    if (lane == 0) b = atomicAdd(queue, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= Bt.B) break;

    // ---- load the instance (osqp_update_lin_cost / osqp_update_bounds scaling)
    double qb_i = (i < n && Bt.q) ? c * (D_i * Bt.q[(size_t)b * n + i]) : 0.0;
    double xi_i = (i < n && S.warm_start && !Bt.fresh) ? Bt.xi[(size_t)b * n + i] : 0.0;
    double lb_r = -1.0, ub_r = 1.0, z_r = 0.0, y_r = 0.0;   // padded rows: A row = 0, never active
    if (r < m) {
      lb_r = E_r * (Bt.l ? Bt.l[(size_t)b * m + r] : P.l0[r]);
      ub_r = E_r * (Bt.u ? Bt.u[(size_t)b * m + r] : P.u0[r]);
      if (S.warm_start && !Bt.fresh) { z_r = Bt.z[(size_t)b * m + r]; y_r = Bt.y[(size_t)b * m + r]; }
    }
    double rho = Bt.fresh ? fmin(fmax(S.rho0, kRhoMin), kRhoMax) : Bt.rho[b];
    int rho_updates = 0;
    // l > u (osqp_update_bounds refuses it) or a row whose class (equality / inequality / free) differs from the
    // shared plan's: the instance is left UNSOLVED (see admm_shared_generic.cu)
    const int ct_now = (lb_r < -kInfty * kMinScaling && ub_r > kInfty * kMinScaling) ? -1 : ((ub_r - lb_r < kRhoTolRow) ? 1 : 0);
    const bool bad_bounds = __any_sync(0xffffffffu, r < m && (lb_r > ub_r || ct_now != ct_r));
    // q̂ = V' q̄
    if (h == 0) sbuf[i] = qb_i;
    __syncwarp();
    double qh_i;
    {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NP / 2; ++k) a = fma(sV[(8 * h + k) * NP + i], sbuf[8 * h + k], a);
      qh_i = a + __shfl_xor_sync(0xffffffffu, a, 16);
    }
    double rv = rho_row(ct_r, rho), rinv = 1.0 / rv;
    double dinv_i = 1.0 / (1.0 + rho * lam_i);
    double dxi_i = 0.0, dy_r = 0.0, t_i = 0.0;
    __syncwarp();
    if (h == 0) cbuf[i] = xi_i;
    cbuf[NP + r] = rv * z_r - y_r;
    __syncwarp();

    int status = SMPC_UNSOLVED, iter = 0;
    bool can_check = false;
    Info I = {};
    double xbar_i = 0.0;

    // OSQP update_info.  cbuf[0..15] holds the current xi.
    auto update_info = [&]() {
      sbuf[r] = y_r;
      __syncwarp();
      double ax = 0.0, apx = 0.0, aty = 0.0, Ax_r = 0.0;
#pragma unroll
      for (int k = 0; k < NP / 2; ++k) {
        double xk = cbuf[8 * h + k];
        ax = fma(sVT[(8 * h + k) * NP + i], xk, ax);
        apx = fma(sPVT[(8 * h + k) * NP + i], xk, apx);
      }
#pragma unroll
      for (int k = 0; k < MP / 2; ++k) aty = fma(sAb[(16 * h + k) * NP + i], sbuf[16 * h + k], aty);
#pragma unroll
      for (int k = 0; k < NP; ++k) Ax_r = fma(wr[k], cbuf[k], Ax_r);
      ax += __shfl_xor_sync(0xffffffffu, ax, 16);
      apx += __shfl_xor_sync(0xffffffffu, apx, 16);
      aty += __shfl_xor_sync(0xffffffffu, aty, 16);
      xbar_i = ax;
      const double rp = Ax_r - z_r, rd = (qb_i + apx) + aty;
      I.s_rp = wmax(fabs(rp)); I.s_z = wmax(fabs(z_r)); I.s_Ax = wmax(fabs(Ax_r));
      I.s_rd = wmax(fabs(rd)); I.s_q = wmax(fabs(qb_i)); I.s_Aty = wmax(fabs(aty)); I.s_Px = wmax(fabs(apx));
      double ob = h == 0 ? 0.5 * ax * apx + qb_i * ax : 0.0;
      if (unscale) {
        I.pri_res = wmax(fabs(Einv_r * rp)); I.nEz = wmax(fabs(Einv_r * z_r)); I.nEAx = wmax(fabs(Einv_r * Ax_r));
        I.dua_res = cinv * wmax(fabs(Dinv_i * rd)); I.nDq = wmax(fabs(Dinv_i * qb_i));
        I.nDAty = wmax(fabs(Dinv_i * aty)); I.nDPx = wmax(fabs(Dinv_i * apx));
        I.obj = cinv * wsum(ob);
      } else {
        I.pri_res = I.s_rp; I.nEz = I.s_z; I.nEAx = I.s_Ax;
        I.dua_res = I.s_rd; I.nDq = I.s_q; I.nDAty = I.s_Aty; I.nDPx = I.s_Px;
        I.obj = wsum(ob);
      }
      if (m == 0) I.pri_res = 0.0;
      __syncwarp();
    };

    auto primal_infeasible = [&](double eps) -> bool {
      double d = dy_r;
      const bool uinf = ub_r > kInfty * kMinScaling, linf = lb_r < -kInfty * kMinScaling;
      if (uinf) d = linf ? 0.0 : fmin(d, 0.0); else if (linf) d = fmax(d, 0.0);
      dy_r = d;
      const double nd = wmax(fabs(unscale ? E_r * d : d));
      if (!(nd > eps)) return false;
      double lhs = 0.0;
      const double dp = fmax(d, 0.0), dm = fmin(d, 0.0);
      if (dp != 0.0) lhs += ub_r * dp;
      if (dm != 0.0) lhs += lb_r * dm;
      lhs = wsum(lhs);
      if (!(lhs < -eps * nd)) return false;
      sbuf[r] = d;
      __syncwarp();
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < MP / 2; ++k) a = fma(sAb[(16 * h + k) * NP + i], sbuf[16 * h + k], a);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      __syncwarp();
      return wmax(fabs(unscale ? Dinv_i * a : a)) < eps * nd;
    };

    auto dual_infeasible = [&](double eps) -> bool {
      if (h == 0) sbuf[i] = dxi_i;
      __syncwarp();
      double dx = 0.0, pd = 0.0, ad = 0.0;
#pragma unroll
      for (int k = 0; k < NP / 2; ++k) {
        double dk = sbuf[8 * h + k];
        dx = fma(sVT[(8 * h + k) * NP + i], dk, dx);
        pd = fma(sPVT[(8 * h + k) * NP + i], dk, pd);
      }
#pragma unroll
      for (int k = 0; k < NP; ++k) ad = fma(wr[k], sbuf[k], ad);
      dx += __shfl_xor_sync(0xffffffffu, dx, 16);
      pd += __shfl_xor_sync(0xffffffffu, pd, 16);
      __syncwarp();
      const double nd = wmax(fabs(unscale ? D_i * dx : dx));
      const double qd = wsum(h == 0 ? qb_i * dx : 0.0);
      const double cs = unscale ? c : 1.0;
      if (!(nd > eps)) return false;
      if (!(qd < -cs * eps * nd)) return false;
      if (!(wmax(fabs(unscale ? Dinv_i * pd : pd)) < cs * eps * nd)) return false;
      if (unscale) ad *= Einv_r;
      const int bad = ((ub_r < kInfty * kMinScaling) && (ad > eps * nd)) || ((lb_r > -kInfty * kMinScaling) && (ad < -eps * nd));
      return !__any_sync(0xffffffffu, bad && r < m);
    };

    auto check_termination = [&](bool approx) -> bool {
      double ea = S.eps_abs, er = S.eps_rel, epi = S.eps_prim_inf, edi = S.eps_dual_inf;
      if (approx) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
      bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
      if (m == 0) prim_ok = true;
      else {
        const double ep = ea + er * fmax(I.nEz, I.nEAx);
        if (I.pri_res < ep) prim_ok = true; else prim_inf = primal_infeasible(epi);
      }
      const double ed = ea + er * (unscale ? cinv : 1.0) * fmax(fmax(I.nDq, I.nDAty), I.nDPx);
      if (I.dua_res < ed) dual_ok = true; else dual_inf = dual_infeasible(edi);
      if (prim_ok && dual_ok) { status = approx ? SMPC_SOLVED_INACCURATE : SMPC_SOLVED; return true; }
      if (prim_inf) { status = approx ? SMPC_PRIMAL_INFEASIBLE_INACCURATE : SMPC_PRIMAL_INFEASIBLE; I.obj = kInfty; return true; }
      if (dual_inf) { status = approx ? SMPC_DUAL_INFEASIBLE_INACCURATE : SMPC_DUAL_INFEASIBLE; I.obj = -kInfty; return true; }
      return false;
    };

    if (!bad_bounds) {
      for (iter = 1; iter <= S.max_iter; ++iter) {
        // t = (sigma G xi + W' w - q̂) ./ (1 + rho lambda): each half-warp sums 24 of the 48 concatenated terms
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const double2 *cv = reinterpret_cast<const double2 *>(cbuf + KH * h);
#pragma unroll
        for (int j = 0; j < KH / 4; ++j) {
          const double2 u0 = cv[2 * j], u1 = cv[2 * j + 1];
          a0 = fma(m1[4 * j + 0], u0.x, a0); a1 = fma(m1[4 * j + 1], u0.y, a1);
          a2 = fma(m1[4 * j + 2], u1.x, a2); a3 = fma(m1[4 * j + 3], u1.y, a3);
        }
        double acc = (a0 + a1) + (a2 + a3);
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        t_i = (acc - qh_i) * dinv_i;
        if (h == 0) tbuf[i] = t_i;
        __syncwarp();
        // z̃ = W t
        const double2 *tv = reinterpret_cast<const double2 *>(tbuf);
        double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
        for (int j = 0; j < NP / 4; ++j) {
          const double2 u0 = tv[2 * j], u1 = tv[2 * j + 1];
          b0 = fma(wr[4 * j + 0], u0.x, b0); b1 = fma(wr[4 * j + 1], u0.y, b1);
          b2 = fma(wr[4 * j + 2], u1.x, b2); b3 = fma(wr[4 * j + 3], u1.y, b3);
        }
        const double zt = (b0 + b1) + (b2 + b3);
        // x, z, y updates (OSQP update_x / update_z / update_y)
        const double xn = alpha * t_i + (1.0 - alpha) * xi_i;
        dxi_i = xn - xi_i; xi_i = xn;
        const double zr = alpha * zt + (1.0 - alpha) * z_r;
        const double zn = fmin(fmax(zr + rinv * y_r, lb_r), ub_r);
        dy_r = rv * (zr - zn);
        z_r = zn; y_r += dy_r;
        can_check = S.check_every && (iter % S.check_every == 0);
        const bool adapt = S.adaptive_rho && S.rho_interval && (iter % S.rho_interval == 0);
        if (can_check || adapt) {
          if (h == 0) cbuf[i] = xi_i;
          __syncwarp();
          update_info();
          if (can_check && check_termination(false)) break;
          if (adapt) {
            const double pr = I.s_rp / (fmax(I.s_z, I.s_Ax) + kDivTol);
            const double dr = I.s_rd / (fmax(fmax(I.s_q, I.s_Aty), I.s_Px) + kDivTol);
            const double rn = fmin(fmax(rho * sqrt(pr / (dr + kDivTol)), kRhoMin), kRhoMax);
            if (rn > rho * S.rho_tol || rn < rho / S.rho_tol) {
              rho = rn; ++rho_updates;
              rv = rho_row(ct_r, rho); rinv = 1.0 / rv; dinv_i = 1.0 / (1.0 + rho * lam_i);
            }
          }
        } else if (h == 0) cbuf[i] = xi_i;
        cbuf[NP + r] = rv * z_r - y_r;
        __syncwarp();
      }
      if (iter > S.max_iter) iter = S.max_iter;
      if (!can_check) {
        if (h == 0) cbuf[i] = xi_i;
        __syncwarp();
        update_info();
        check_termination(false);
      }
      if (status == SMPC_UNSOLVED) { if (!check_termination(true)) status = SMPC_MAX_ITER_REACHED; }
    }
    __syncwarp();

    // ---- store_solution
    const bool has_sol = !bad_bounds && !(status == SMPC_PRIMAL_INFEASIBLE || status == SMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                                          status == SMPC_DUAL_INFEASIBLE || status == SMPC_DUAL_INFEASIBLE_INACCURATE);
    if (h == 0 && i < n) {
      if (Bt.x_out) Bt.x_out[(size_t)b * n + i] = has_sol ? D_i * xbar_i : qnan;
      Bt.xi[(size_t)b * n + i] = has_sol ? xi_i : 0.0;
    }
    if (r < m) {
      if (Bt.y_out) Bt.y_out[(size_t)b * m + r] = has_sol ? cinv * (E_r * y_r) : qnan;
      Bt.z[(size_t)b * m + r] = has_sol ? z_r : 0.0;
      Bt.y[(size_t)b * m + r] = has_sol ? y_r : 0.0;
    }
    if (lane == 0) {
      Bt.rho[b] = rho;
      Bt.status[b] = status; Bt.iter[b] = iter; Bt.rho_updates[b] = rho_updates;
      Bt.obj[b] = I.obj; Bt.pri_res[b] = I.pri_res; Bt.dua_res[b] = I.dua_res;
    }
  }
}

bool small_kernel_supports(int n, int m) { return n >= 1 && n <= NP && m >= 0 && m <= MP; }

size_t small_pack_doubles() { return (size_t)(NP + MP) * NP + NP * MP + 3 * NP * NP + MP * NP + 3 * NP + 2 * MP; }

cudaError_t launch_admm_shared_small(const SmallPackDev &K, const SharedPlanDev &P, const BatchDev &Bt,
                                     const SettingsDev &S, int *queue, int num_sms, cudaStream_t stream) {
  const int wpc = 4;
  const size_t smem = (size_t)(kCtaMatDoubles + wpc * kWarpDoubles) * sizeof(double);
  cudaError_t e = cudaMemsetAsync(queue, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  int grid = (Bt.B + wpc - 1) / wpc;
  const int resident = num_sms * 4;   // __launch_bounds__(128, 3): four CTAs per SM
  if (grid > resident) grid = resident;
  admm_shared_small_kernel<<<grid, wpc * 32, smem, stream>>>(K, P, Bt, S, queue);
  return cudaGetLastError();
}

}  // namespace smpc
