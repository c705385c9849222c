// admm_shared_small_fused.cu -- one-phase register-resident ADMM kernel for small QPs whose rows are pairs [G; -G]
// (n <= 16, m = 2 mp <= 32: the reference's own problem, n = 15, m = 30, cpp:335).
//
// admm_shared_small_kernel<true> runs OSQP's iteration as two dependent mat-vecs, t = dinv .* (M1 s - q̂) and z̃_top = W_top t
// (M1 = [sigma G | W_top'], s = [xi; wd], wd = w_top - w_bot), each ending in a shuffle reduction and a shared-memory broadcast:
// 392 cycles per iteration alone on an SM.  Here t is substituted:
//     [t; z̃_top] = K(rho) s + k0(rho),    K = [diag(dinv) M1; W_top diag(dinv) M1] (32 x 32),   dinv = 1 / (1 + rho lambda)
// so an iteration is ONE mat-vec, one shared-memory round trip and two shuffle levels (243 cycles alone,
// profiles/microbench/fused_row_iteration.cu).  K depends on the instance's rho, so it lives in registers per instance
// (64 registers) and is rebuilt when rho adapts (256 DFMA per lane, once or twice per solve); K(rho0) is built once per CTA in
// shared memory, so cold solves start without a rebuild.
// Layout: lane (a, b) = (lane >> 2, lane & 3) holds a 4 x 8 block of K -- four rows of row group a, columns 8b .. 8b+7 -- and
// reads only ITS eight entries of s (four LDS.128, four distinct 16-byte chunks per instruction, stored contiguously: no bank
// conflict and a quarter of the shared-memory traffic of a full broadcast).  The four partial sums are transpose-reduced over
// the four b-lanes with three 64-bit shuffles; the local row order of a lane (small_rowmix) makes the kept / sent halves
// compile-time.  After the reduction lane (a, b) owns one output row: t_idx on the "xi-lanes" (b < 2) or z̃_top,idx on the
// "pair-lanes" (b >= 2), idx = 2a + (b & 1).  The element-wise update is the same code on every lane: a pair-lane updates
// BOTH rows of its pair (so no exchange shuffle), a xi-lane is the degenerate pair (bounds -+inf, rho_vec = (1, 0), alpha_b = 0)
// for which the same expressions yield xi' = alpha t + (1 - alpha) xi.
// Termination checks (update_info, check_termination, both infeasibility tests, adapt_rho) use the same block layout:
// two 32 x 16 passes, [V; W_top] xi and [P̄V xi; A̅_top' (y_top - y_bot)], operators from shared memory as LDS.128 in a
// [chunk][lane] layout, inlined at ONE call site (K stays in registers): ~700 cycles per check instead of ~2 850.
// Iterates follow the oracle (oracle/osqp_port.c) to round-off like the two-phase kernels; queue, difficulty classes and
// quiet SMs are those of admm_shared_small.cu.  Opt-in while it is being tuned: SMPC_SMALL_FUSED=1 (environment).
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "classify.cuh"
#include "device_types.cuh"
#include "kernels.cuh"
#include "small_common.cuh"

namespace smpc {

namespace {

// CTA-shared block (doubles)
constexpr int oM1x = 0;                  // [16 k][32]   [sigma G | W_top'](k, .) in small_pos32 order
constexpr int oWTt = oM1x + 512;         // [16 k][16 i] W_top(i, k)
constexpr int oC1 = oWTt + 256;          // [8 chunks][32 lanes][2]
constexpr int oC2 = oC1 + 512;
constexpr int oV = oC2 + 512;            // [16][16] V row-major (q̂ = V' q̄)
constexpr int oCst = oV + 256;           // [32][4] per-lane scalings
constexpr int oK0 = oCst + 128;          // [32 j][32 lanes] K(rho0) blocks, j = 8 l + c
constexpr int oWD0 = oK0 + 1024;         // [16 k][32 lanes] W_top(idx(lane), k) dinv_k(rho0)
constexpr int oDv0 = oWD0 + 512;         // [16] dinv(rho0)
constexpr int kFusedCtaDoubles = oDv0 + 16;
// per-warp block (doubles): s[32] | cv = [xi-part 16 | pair-part 16] | qh[16] | dv[16]
constexpr int kFusedWarpDoubles = 32 + 32 + 16 + 16;

__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
// transpose-reduction of four partial sums over the four b-lanes of a row group: returns the lane's own row
__device__ __forceinline__ double treduce4(double p0, double p1, double p2, double p3) {
  const double q0 = p0 + __shfl_xor_sync(kFull, p2, 1), q1 = p1 + __shfl_xor_sync(kFull, p3, 1);
  return q0 + __shfl_xor_sync(kFull, q1, 2);
}
// One 32-row x 16-column product of a check.  a_op: this lane's 16-byte slot of the operator pack ([chunk][lane] layout,
// chunk stride 512 B); a_even / a_odd: the lane's four columns (two chunks, 64 B apart) of the input vector of its own-kind
// rows (local rows 0, 2) and of the other kind's rows (local rows 1, 3).
__device__ __forceinline__ double check_pass(uint32_t a_op, uint32_t a_even, uint32_t a_odd) {
  const double2 e0 = lds128(a_even), e1 = lds128(a_even + 64), o0 = lds128(a_odd), o1 = lds128(a_odd + 64);
  double p[4];
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const double2 c0 = lds128(a_op + 1024 * l), c1 = lds128(a_op + 1024 * l + 512);
    const double2 u0 = (l & 1) ? o0 : e0, u1 = (l & 1) ? o1 : e1;
    p[l] = fma(c0.x, u0.x, c0.y * u0.y) + fma(c1.x, u1.x, c1.y * u1.y);
  }
  return treduce4(p[0], p[1], p[2], p[3]);
}

struct FusedCheck {
  double rho, obj, pri_res, dua_res, xbar;
  int status, rho_changed;
};

// OSQP update_info + check_termination (+ is_primal_infeasible / is_dual_infeasible) + adapt_rho for one QP in the lane roles
// of the fused kernel.  xi-lanes pass xi in z_t, delta_xi in dz_t and q̄_idx in qb; pair-lanes pass z, y, delta_y and the bounds
// of both rows of their pair (xi-lanes: z_b = y_t = y_b = dy_t = dy_b = 0, bounds (-inf, inf) / (-1, 1)).  sc = the lane's
// scalings (oCst).  cv (a_cvme / a_cvxi / a_cvyd) is scratch.  Everything returned except xbar is uniform across the warp.
__device__ __forceinline__ FusedCheck fused_check(const SettingsDev &S, bool is_xi, bool pair_ok, uint32_t a_c1, uint32_t a_c2, uint32_t a_cvme,
                                                  uint32_t a_cvxi, uint32_t a_cvyd, double c, double cinv, double2 sc0, double2 sc1,
                                                  double rho, double qb, double s_q, double nDq, double z_t, double z_b, double y_t,
                                                  double y_b, double dy_t, double dy_b, double dz_t, double lb_t, double ub_t,
                                                  double lb_b, double ub_b, bool do_check, bool approx, bool do_adapt, bool want_obj) {
  const bool unscale = !S.scaled_termination;
  const uint32_t a_ev2 = is_xi ? a_cvxi : a_cvyd, a_od2 = is_xi ? a_cvyd : a_cvxi;
  // xi-lanes: sc0 = (D, 1/D); pair-lanes: sc0 = (E, 1/E) of the top row, sc1 = (E, 1/E) of the bottom row
  const double Dn = sc0.x, Dinv = sc0.y, E_t = sc0.x, Einv_t = sc0.y, E_b = sc1.x, Einv_b = sc1.y;
  FusedCheck o;
  o.rho = rho; o.status = SMPC_UNSOLVED; o.rho_changed = 0; o.obj = 0.0;
  sts64(a_cvme, is_xi ? z_t : y_t - y_b);
  __syncwarp();
  const double o1 = check_pass(a_c1, a_cvxi, a_cvxi);    // xi-lane: x̄_idx ; pair-lane: (A̅ x̄)_top,idx = (W_top xi)_idx
  const double o2 = check_pass(a_c2, a_ev2, a_od2);      // xi-lane: (P̄ x̄)_idx ; pair-lane: (A̅' y)_idx
  const double o2x = __shfl_xor_sync(kFull, o2, 2);      // ... which belongs on the xi-lane of the same idx
  __syncwarp();
  o.xbar = o1;
  // n-space quantities live on the xi-lanes, m-space quantities on the pair-lanes; the other kind contributes 0 to a norm
  const double apx = o2, aty = o2x, rd = (qb + apx) + aty;                    // (xi-lanes)
  const double rp_t = o1 - z_t, rp_b = -o1 - z_b;                             // (pair-lanes)
  const ull kx = is_xi ? ~0ULL : 0ULL, kp = ~kx;
  const double s_rp = wmax_bits(umax2(abits(rp_t), abits(rp_b)) & kp), s_z = wmax_bits(umax2(abits(z_t), abits(z_b)) & kp);
  const double s_Ax = wmax_bits(abits(o1) & kp);
  const double s_rd = wmax_bits(abits(rd) & kx), s_Aty = wmax_bits(abits(aty) & kx), s_Px = wmax_bits(abits(apx) & kx);
  double pri_res, dua_res, nEz, nEAx, nDAty, nDPx, nq;
  if (unscale) {
    pri_res = wmax_bits(umax2(abits(Einv_t * rp_t), abits(Einv_b * rp_b)) & kp);
    nEz = wmax_bits(umax2(abits(Einv_t * z_t), abits(Einv_b * z_b)) & kp);
    nEAx = wmax_bits(umax2(abits(Einv_t * o1), abits(Einv_b * o1)) & kp);
    dua_res = cinv * wmax_bits(abits(Dinv * rd) & kx);
    nDAty = wmax_bits(abits(Dinv * aty) & kx); nDPx = wmax_bits(abits(Dinv * apx) & kx); nq = nDq;
  } else {
    pri_res = s_rp; nEz = s_z; nEAx = s_Ax; dua_res = s_rd; nDAty = s_Aty; nDPx = s_Px; nq = s_q;
  }
  o.pri_res = pri_res; o.dua_res = dua_res;

  if (do_check) {
    double ea = S.eps_abs, er = S.eps_rel, epi = S.eps_prim_inf, edi = S.eps_dual_inf;
    if (approx) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (pri_res < ea + er * fmax(nEz, nEAx)) prim_ok = true;
    else {
      // is_primal_infeasible: project delta_y on the polar of the recession cone of [l, u], row by row
      double d_t = dy_t, d_b = dy_b;
      {
        const bool uinf = ub_t > kInfty * kMinScaling, linf = lb_t < -kInfty * kMinScaling;
        if (uinf) d_t = linf ? 0.0 : fmin(d_t, 0.0); else if (linf) d_t = fmax(d_t, 0.0);
      }
      {
        const bool uinf = ub_b > kInfty * kMinScaling, linf = lb_b < -kInfty * kMinScaling;
        if (uinf) d_b = linf ? 0.0 : fmin(d_b, 0.0); else if (linf) d_b = fmax(d_b, 0.0);
      }
      const double nd = wmax_bits((unscale ? umax2(abits(E_t * d_t), abits(E_b * d_b)) : umax2(abits(d_t), abits(d_b))) & kp);
      if (nd > epi) {
        double lhs = 0.0;
        const double pt = fmax(d_t, 0.0), mt = fmin(d_t, 0.0), pb = fmax(d_b, 0.0), mb = fmin(d_b, 0.0);
        if (pt != 0.0) lhs += ub_t * pt;
        if (mt != 0.0) lhs += lb_t * mt;
        if (pb != 0.0) lhs += ub_b * pb;
        if (mb != 0.0) lhs += lb_b * mb;
        lhs = wsum(lhs);
        if (lhs < -epi * nd) {
          sts64(a_cvme, is_xi ? z_t : d_t - d_b);
          __syncwarp();
          const double at = __shfl_xor_sync(kFull, check_pass(a_c2, a_ev2, a_od2), 2);   // (A̅' d)_idx on the xi-lanes
          __syncwarp();
          prim_inf = wmax_bits(abits(unscale ? Dinv * at : at) & kx) < epi * nd;
        }
      }
    }
    if (dua_res < ea + er * (unscale ? cinv : 1.0) * fmax(fmax(nq, nDAty), nDPx)) dual_ok = true;
    else {
      // is_dual_infeasible on delta_x = V delta_xi (OSQP's three conditions in its order, each only when the one before holds)
      sts64(a_cvme, is_xi ? dz_t : 0.0);
      __syncwarp();
      const double r1 = check_pass(a_c1, a_cvxi, a_cvxi);   // xi-lane: delta_x_idx ; pair-lane: (A̅ delta_x)_top,idx
      const double dx = is_xi ? r1 : 0.0;
      const double nd = wmax_bits(abits(unscale ? Dn * dx : dx));
      const double cs = unscale ? c : 1.0;
      if (nd > edi && wsum(qb * dx) < -cs * edi * nd) {
        const double r2 = check_pass(a_c2, a_ev2, a_od2);   // xi-lane: (P̄ delta_x)_idx
        const double pd = is_xi ? r2 : 0.0;
        if (wmax_bits(abits(unscale ? Dinv * pd : pd)) < cs * edi * nd) {
          double ad_t = r1, ad_b = -r1;
          if (unscale) { ad_t *= Einv_t; ad_b *= Einv_b; }
          const bool bad_t = ((ub_t < kInfty * kMinScaling) && (ad_t > edi * nd)) || ((lb_t > -kInfty * kMinScaling) && (ad_t < -edi * nd));
          const bool bad_b = ((ub_b < kInfty * kMinScaling) && (ad_b > edi * nd)) || ((lb_b > -kInfty * kMinScaling) && (ad_b < -edi * nd));
          dual_inf = !__any_sync(kFull, pair_ok && (bad_t || bad_b));
        }
      }
      __syncwarp();
    }
    if (prim_ok && dual_ok) o.status = approx ? SMPC_SOLVED_INACCURATE : SMPC_SOLVED;
    else if (prim_inf) o.status = approx ? SMPC_PRIMAL_INFEASIBLE_INACCURATE : SMPC_PRIMAL_INFEASIBLE;
    else if (dual_inf) o.status = approx ? SMPC_DUAL_INFEASIBLE_INACCURATE : SMPC_DUAL_INFEASIBLE;
  }
  // the objective is reported, never tested: only when the solve ends here (or the caller is at max_iter)
  if (o.status != SMPC_UNSOLVED || want_obj) {
    const double ob = wsum(is_xi ? 0.5 * o1 * apx + qb * o1 : 0.0);
    const bool pinf = o.status == SMPC_PRIMAL_INFEASIBLE || o.status == SMPC_PRIMAL_INFEASIBLE_INACCURATE;
    const bool dinf = o.status == SMPC_DUAL_INFEASIBLE || o.status == SMPC_DUAL_INFEASIBLE_INACCURATE;
    o.obj = pinf ? kInfty : (dinf ? -kInfty : (unscale ? cinv * ob : ob));
  }
  if (do_adapt && o.status == SMPC_UNSOLVED) {
    // compute_rho_estimate / adapt_rho on the SCALED residual norms
    const double pr = fast_div(s_rp, fmax(s_z, s_Ax) + kDivTol);
    const double dr = fast_div(s_rd, fmax(fmax(s_q, s_Aty), s_Px) + kDivTol);
    const double rn = fmin(fmax(rho * sqrt(fast_div(pr, dr + kDivTol)), kRhoMin), kRhoMax);
    if (rn > rho * S.rho_tol || rn < rho / S.rho_tol) { o.rho = rn; o.rho_changed = 1; }
  }
  return o;
}

}  // namespace

#ifndef SMPC_FUSED_CTAS
#define SMPC_FUSED_CTAS 3
#endif
constexpr int kFusedCtas = SMPC_FUSED_CTAS;   // resident CTAs per SM the register budget is set for
__global__ void __launch_bounds__(128, kFusedCtas)
admm_shared_small_fused_kernel(SmallPackDev K, SharedPlanDev P, BatchDev Bt, SettingsDev S, int *queue, const int *lists, int quiet_cap) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_rank, s_ticket, s_cnt[kClasses];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a = lane >> 2, b = lane & 3;
  const bool is_xi = b < 2;
  const int idx = 2 * a + (b & 1), idx1 = idx ^ 1;
  const int n = P.n, m = P.m, mp = m >> 1;
  // quiet-SM scheduling: see admm_shared_small_kernel
  if (quiet_cap > 0 && tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smid &= kSmSlots - 1;
    const int rank = atomicAdd(queue + kQRank + smid, 1);
    int t;
    if (rank == 0) { t = atomicAdd(queue + kQSeen, 1); atomicExch(queue + kQTicket + smid, t + 1); }
    else { do { t = atomicAdd(queue + kQTicket + smid, 0); } while (t == 0); t -= 1; }
    s_rank = rank; s_ticket = t;
  }
  // ---- constant operators, K(rho0)
  for (int e = tid; e < 512; e += blockDim.x) { smem[oM1x + e] = K.M1x[e]; smem[oC1 + e] = K.C1[e]; smem[oC2 + e] = K.C2[e]; }
  for (int e = tid; e < 256; e += blockDim.x) { smem[oWTt + e] = K.WTt[e]; smem[oV + e] = K.V[e]; }
  if (tid < 128) smem[oCst + tid] = K.cst[tid];
  const double rho0 = fmin(fmax(S.rho0, kRhoMin), kRhoMax);
  if (tid < 16) smem[oDv0 + tid] = 1.0 / (1.0 + rho0 * K.lam[tid]);
  __syncthreads();
  for (int e = tid; e < 1024; e += blockDim.x) {
    const int ln = e & 31, j = e >> 5, l = j >> 3, cc = j & 7, a2 = ln >> 2, b2 = ln & 3;
    const int bt = b2 ^ small_rowmix(l), ri = 2 * a2 + (bt & 1), col = small_pos32(8 * b2 + cc);
    double val;
    if (bt < 2) val = smem[oDv0 + ri] * smem[oM1x + ri * 32 + col];
    else {
      val = 0.0;
      for (int k = 0; k < 16; ++k) val = fma(smem[oWTt + k * 16 + ri] * smem[oDv0 + k], smem[oM1x + k * 32 + col], val);
    }
    smem[oK0 + e] = val;
  }
  for (int e = tid; e < 512; e += blockDim.x) {
    const int ln = e & 31, k = e >> 5;
    smem[oWD0 + e] = smem[oWTt + k * 16 + 2 * (ln >> 2) + (ln & 1)] * smem[oDv0 + k];
  }

  double *wbuf = smem + kFusedCtaDoubles + warp * kFusedWarpDoubles;
  double *sv = wbuf, *cv = wbuf + 32, *qh = wbuf + 64, *dv = wbuf + 80;
  const uint32_t a_s = (uint32_t)__cvta_generic_to_shared(sv) + 16 * b;
  const uint32_t a_sme = (uint32_t)__cvta_generic_to_shared(sv) + 8 * small_pos32(is_xi ? idx : 16 + idx);
  const uint32_t a_cvxi = (uint32_t)__cvta_generic_to_shared(cv) + 16 * b, a_cvyd = a_cvxi + 128;
  const uint32_t a_cvme = (uint32_t)__cvta_generic_to_shared(cv) + 8 * ((is_xi ? 0 : 16) + small_pos16(idx));
  const uint32_t a_c1 = (uint32_t)__cvta_generic_to_shared(smem + oC1) + 16 * lane, a_c2 = a_c1 + 8 * 512;
  const uint32_t a_m1 = (uint32_t)__cvta_generic_to_shared(smem + oM1x) + 16 * b;
  const uint32_t a_cst = (uint32_t)__cvta_generic_to_shared(smem + oCst) + 32 * lane;
  const bool pair_ok = !is_xi && idx < mp;
  const int ct_t = pair_ok ? K.ctype[idx] : 0, ct_b = pair_ok ? K.ctype[mp + idx] : 0;
  const double alpha = S.alpha, oma = 1.0 - S.alpha, alpha_b = is_xi ? 0.0 : -S.alpha, c = P.c, cinv = P.cinv;
  const double inf = __longlong_as_double(0x7ff0000000000000LL), qnan = __longlong_as_double(0x7ff8000000000000LL);
  const int check_every = S.check_every > 0 ? S.check_every : 0x7fffffff;
  const int adapt_every = (S.adaptive_rho && S.rho_interval > 0) ? S.rho_interval : 0x7fffffff;
  __syncthreads();
  // programmatic dependent launch: everything above reads only the constant packs (see admm_shared_small_kernel)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // the class sizes do not change during the solve: one read per CTA instead of up to five dependent reads per instance
  if (tid < kClasses) s_cnt[tid] = lists ? queue[1 + tid] : 0;
  __syncthreads();
  int n_quiet = 0;
  bool quiet_warp = false;
  if (quiet_cap > 0 && lists != nullptr) {
    n_quiet = min(s_cnt[0], 4 * quiet_cap);
    const int sq = (n_quiet + 3) >> 2;
    quiet_warp = s_ticket < sq && s_rank == 0;
    if (s_ticket < sq && s_rank != 0) { release_queue_warp(queue, gridDim.x * (blockDim.x >> 5), lane); return; }
  }

  for (;;) {
    int bi = 0;
    if (lane == 0) {
      bi = -1;
      if (quiet_warp) {
        const int qh_ = atomicAdd(queue + kQHard, 1);
        SMPC_DBG(qh_ >= 0 && n_quiet <= Bt.B, "quiet-share counter");
        if (qh_ < n_quiet) bi = lists[qh_]; else quiet_warp = false;
      }
      if (bi < 0) {
        int q = atomicAdd(queue, 1);
        if (q < Bt.B - n_quiet) {
          if (lists == nullptr) bi = q;
          else {
#pragma unroll
            for (int k = 0; k < kClasses; ++k) {
              const int skip = k == 0 ? n_quiet : 0, cnt = s_cnt[k] - skip;
              SMPC_DBG(cnt >= 0 && cnt + skip <= Bt.B, "class size");
              if (bi < 0) { if (q < cnt) bi = lists[(size_t)k * Bt.B + skip + q]; else q -= cnt; }
            }
          }
        }
      }
    }
    SMPC_DBG(bi < Bt.B, "instance index from the queue");
    bi = __shfl_sync(kFull, bi, 0);
    if (bi < 0) break;

    // ---- load the instance (osqp_update_lin_cost / osqp_update_bounds scaling)
    const bool warm = S.warm_start && !Bt.fresh;
    double qb = 0.0, z_t = 0.0, z_b = 0.0, y_t = 0.0, y_b = 0.0;
    double lb_t = is_xi ? -inf : -1.0, ub_t = is_xi ? inf : 1.0, lb_b = -1.0, ub_b = 1.0;   // padded pairs: A row = 0, never active
    {
      const double2 sc0 = lds128(a_cst), sc1 = lds128(a_cst + 16);
      if (is_xi) {
        if (idx < n) {
          if (Bt.q) qb = c * (sc0.x * Bt.q[(size_t)bi * n + idx]);
          if (warm) z_t = Bt.xi[(size_t)bi * n + idx];
        }
      } else if (pair_ok) {
        const size_t rt = (size_t)bi * m + idx, rb = rt + mp;
        lb_t = sc0.x * (Bt.l ? Bt.l[rt] : P.l0[idx]); ub_t = sc0.x * (Bt.u ? Bt.u[rt] : P.u0[idx]);
        lb_b = sc1.x * (Bt.l ? Bt.l[rb] : P.l0[mp + idx]); ub_b = sc1.x * (Bt.u ? Bt.u[rb] : P.u0[mp + idx]);
        if (warm) { z_t = Bt.z[rt]; z_b = Bt.z[rb]; y_t = Bt.y[rt]; y_b = Bt.y[rb]; }
      }
    }
    double rho = Bt.fresh ? rho0 : Bt.rho[bi];
    int rho_updates = 0;
    // l > u (osqp_update_bounds refuses it): left UNSOLVED; a row whose class differs from the plan's keeps the plan's rho_vec entry
    const bool bad_row = pair_ok && (lb_t > ub_t || lb_b > ub_b);
    const bool bad_bounds = __any_sync(kFull, bad_row);
    // q̂ = V' q̄ (the summation order of the two-phase kernel), norms of q̄ for the dual tolerance
    if (is_xi) dv[idx] = qb;
    __syncwarp();
    {
      const int h = lane >> 4, i = lane & 15;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < NP / 2; ++k) acc = fma(smem[oV + (8 * h + k) * NP + i], dv[8 * h + k], acc);
      acc += __shfl_xor_sync(kFull, acc, 16);
      __syncwarp();
      if (h == 0) qh[i] = acc;
    }
    const double s_q = wmax_bits(abits(qb));
    double nDq;
    { const double2 sc0 = lds128(a_cst); nDq = wmax_bits(is_xi ? abits(sc0.y * qb) : 0ULL); }
    __syncwarp();

    int status = SMPC_UNSOLVED, iter = 0;
    int to_check = check_every, to_adapt = adapt_every;
    double r_obj = 0.0, r_pri = 0.0, r_dua = 0.0, r_xbar = 0.0;
    double dy_t = 0.0, dy_b = 0.0, dz_t = 0.0;

    if (!bad_bounds) {
      double kr[4][8], k0 = 0.0;
      double rv_t = 1.0, rv_b = 0.0, base_t = 0.0, base_b = 0.0, d_t = 0.0, d_b = 0.0;
      bool need_k = true;
      for (;;) {
        if (need_k) {
          // ---- everything that depends on rho: rho_vec of the lane's rows, the carried bases, s, K(rho) and k0
          need_k = false;
          rv_t = is_xi ? 1.0 : rho_row(ct_t, rho); rv_b = is_xi ? 0.0 : rho_row(ct_b, rho);
          const double ri_t = 1.0 / rv_t, ri_b = is_xi ? 0.0 : 1.0 / rv_b;
          base_t = fma(ri_t, y_t, oma * z_t); base_b = fma(ri_b, y_b, oma * z_b);
          sts64(a_sme, (rv_t * z_t - y_t) - (rv_b * z_b - y_b));
          if (rho == rho0) {
#pragma unroll
            for (int l = 0; l < 4; ++l)
#pragma unroll
              for (int cc = 0; cc < 8; ++cc) kr[l][cc] = smem[oK0 + (8 * l + cc) * 32 + lane];
            if (is_xi) k0 = -(smem[oDv0 + idx] * qh[idx]);
            else {
              double acc = 0.0;
#pragma unroll
              for (int k = 0; k < 16; ++k) acc = fma(smem[oWD0 + k * 32 + lane], qh[k], acc);
              k0 = -acc;
            }
          } else {
            if (is_xi) { const double2 sc1 = lds128(a_cst + 16); dv[idx] = 1.0 / (1.0 + rho * sc1.x); }
            __syncwarp();
            double Z0[8], Z1[8], kz = 0.0;
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) { Z0[cc] = 0.0; Z1[cc] = 0.0; }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const double dk = dv[k], c0 = smem[oWTt + k * 16 + idx] * dk, c1 = smem[oWTt + k * 16 + idx1] * dk;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const double2 mm = lds128(a_m1 + 256 * k + 64 * j);
                Z0[2 * j] = fma(c0, mm.x, Z0[2 * j]); Z0[2 * j + 1] = fma(c0, mm.y, Z0[2 * j + 1]);
                Z1[2 * j] = fma(c1, mm.x, Z1[2 * j]); Z1[2 * j + 1] = fma(c1, mm.y, Z1[2 * j + 1]);
              }
              kz = fma(c0, qh[k], kz);
            }
            const double d0 = dv[idx], d1 = dv[idx1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const double2 m0 = lds128(a_m1 + 256 * idx + 64 * j), m1 = lds128(a_m1 + 256 * idx1 + 64 * j);
              const double t0x = d0 * m0.x, t0y = d0 * m0.y, t1x = d1 * m1.x, t1y = d1 * m1.y;
              // local rows 0 / 2: own kind (t on xi-lanes, z̃ on pair-lanes), rows idx / idx ^ 1; local rows 1 / 3: the other kind
              kr[0][2 * j] = is_xi ? t0x : Z0[2 * j]; kr[0][2 * j + 1] = is_xi ? t0y : Z0[2 * j + 1];
              kr[1][2 * j] = is_xi ? Z0[2 * j] : t0x; kr[1][2 * j + 1] = is_xi ? Z0[2 * j + 1] : t0y;
              kr[2][2 * j] = is_xi ? t1x : Z1[2 * j]; kr[2][2 * j + 1] = is_xi ? t1y : Z1[2 * j + 1];
              kr[3][2 * j] = is_xi ? Z1[2 * j] : t1x; kr[3][2 * j + 1] = is_xi ? Z1[2 * j + 1] : t1y;
            }
            k0 = is_xi ? -(d0 * qh[idx]) : -kz;
          }
          __syncwarp();
        }
        // ---- iterations until the next termination check / rho adaptation
        int steps = to_check < to_adapt ? to_check : to_adapt;
        if (steps > S.max_iter - iter) steps = S.max_iter - iter;
        // one iteration: [t; z̃_top] = K s + k0, then the z, y updates (OSQP update_z / update_y re-associated as in
        // admm_shared_small.cu): v = alpha z̃ + base, z = clip(v), d = v - z (y = rho d), w' = rho (2z - v), base' = (1 - alpha) z + d;
        // xi-lanes: z = v = alpha t + (1 - alpha) xi
        auto iteration = [&]() {
          const double2 u0 = lds128(a_s), u1 = lds128(a_s + 64), u2 = lds128(a_s + 128), u3 = lds128(a_s + 192);
          double p[4];
#pragma unroll
          for (int l = 0; l < 4; ++l) {
            double e0 = l == 0 ? k0 : 0.0, e1 = 0.0;
            e0 = fma(kr[l][0], u0.x, e0); e1 = fma(kr[l][1], u0.y, e1);
            e0 = fma(kr[l][2], u1.x, e0); e1 = fma(kr[l][3], u1.y, e1);
            e0 = fma(kr[l][4], u2.x, e0); e1 = fma(kr[l][5], u2.y, e1);
            e0 = fma(kr[l][6], u3.x, e0); e1 = fma(kr[l][7], u3.y, e1);
            p[l] = e0 + e1;
          }
          const double acc = treduce4(p[0], p[1], p[2], p[3]);
          const double vt = fma(alpha, acc, base_t), vb = fma(alpha_b, acc, base_b);
          const double zt = clip_sel(vt, lb_t, ub_t), zb = clip_sel(vb, lb_b, ub_b);
          d_t = vt - zt; d_b = vb - zb;
          __syncwarp();
          sts64(a_sme, fma(rv_t, fma(2.0, zt, -vt), -(rv_b * fma(2.0, zb, -vb))));
          base_t = fma(oma, zt, d_t); base_b = fma(oma, zb, d_b);
          z_t = zt; z_b = zb;
          __syncwarp();
        };
        for (int s = 1; s < steps; ++s) iteration();
        const double zp = z_t, dtp = d_t, dbp = d_b;   // state before the last iteration of the stretch
        if (steps > 0) iteration();
        if (steps > 0) {
          // delta_xi, delta_y and y of the last iteration (y is carried as d = v - z in between)
          const double yn_t = rv_t * d_t, yn_b = rv_b * d_b;
          dy_t = yn_t - (steps > 1 ? rv_t * dtp : y_t); dy_b = yn_b - (steps > 1 ? rv_b * dbp : y_b);
          y_t = yn_t; y_b = yn_b;
          dz_t = z_t - zp;
        }
        iter += steps; to_check -= steps; to_adapt -= steps;
        const bool at_max = iter >= S.max_iter;
        const bool sched_check = to_check == 0, sched_adapt = to_adapt == 0;
        if (sched_check) to_check = check_every;
        if (sched_adapt) to_adapt = adapt_every;
        // passes over the check routine: 0 = the scheduled one (check and / or rho adaptation), 1 = the exact check OSQP makes at
        // max_iter when the last iteration was not a check, 2 = the approximate check (10 x tolerances) after it
        int pass = (sched_check || sched_adapt) ? 0 : 1;
        for (;;) {
          const double2 sc0 = lds128(a_cst), sc1 = lds128(a_cst + 16);
          const FusedCheck co = fused_check(S, is_xi, pair_ok, a_c1, a_c2, a_cvme, a_cvxi, a_cvyd, c, cinv, sc0, sc1, rho, qb, s_q, nDq,
                                            z_t, z_b, y_t, y_b, dy_t, dy_b, dz_t, lb_t, ub_t, lb_b, ub_b,
                                            pass == 0 ? sched_check : true, pass == 2, pass == 0 && sched_adapt, at_max);
          r_pri = co.pri_res; r_dua = co.dua_res; r_xbar = co.xbar;
          if (co.status != SMPC_UNSOLVED || at_max) r_obj = co.obj;
          if (co.status != SMPC_UNSOLVED) { status = co.status; break; }
          if (co.rho_changed) { rho = co.rho; ++rho_updates; need_k = true; }
          if (!at_max) break;
          if (pass == 0 && !sched_check) pass = 1;
          else if (pass < 2) pass = 2;
          else { status = SMPC_MAX_ITER_REACHED; break; }
        }
        if (status != SMPC_UNSOLVED) break;
      }
    }
    __syncwarp();

    // ---- store_solution
    const bool has_sol = !bad_bounds && !(status == SMPC_PRIMAL_INFEASIBLE || status == SMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                                          status == SMPC_DUAL_INFEASIBLE || status == SMPC_DUAL_INFEASIBLE_INACCURATE);
    const double2 sc0 = lds128(a_cst), sc1 = lds128(a_cst + 16);
    if (is_xi) {
      if (idx < n) {
        if (Bt.x_out) Bt.x_out[(size_t)bi * n + idx] = has_sol ? sc0.x * r_xbar : qnan;
        Bt.xi[(size_t)bi * n + idx] = has_sol ? z_t : 0.0;
      }
    } else if (pair_ok) {
      const size_t rt = (size_t)bi * m + idx, rb = rt + mp;
      if (Bt.y_out) { Bt.y_out[rt] = has_sol ? cinv * (sc0.x * y_t) : qnan; Bt.y_out[rb] = has_sol ? cinv * (sc1.x * y_b) : qnan; }
      Bt.z[rt] = has_sol ? z_t : 0.0; Bt.z[rb] = has_sol ? z_b : 0.0;
      Bt.y[rt] = has_sol ? y_t : 0.0; Bt.y[rb] = has_sol ? y_b : 0.0;
    }
    if (lane == 0) {   // xi-lane of entry 0
      Bt.rho[bi] = rho;
      Bt.status[bi] = status; Bt.iter[bi] = iter; Bt.rho_updates[bi] = rho_updates;
      Bt.obj[bi] = r_obj; Bt.pri_res[bi] = r_pri; Bt.dua_res[bi] = r_dua;
      if (Bt.u_apply) {
        double un = 0.0;
        if (status == SMPC_SOLVED || Bt.u_export) un = Bt.u_apply[bi];
        if (status == SMPC_SOLVED) { un = __dadd_rn(un, __dmul_rn(sc0.x, r_xbar)); Bt.u_apply[bi] = un; }   // U += dU*[0] (cpp:105): x[0] rounded first, no FMA
        if (Bt.u_export) Bt.u_export[bi] = un;
      }
      if (Bt.status_export) Bt.status_export[bi] = status;
    }
  }
  // the last warp of the grid to leave re-arms the queue for the next launch
  release_queue_warp(queue, gridDim.x * (blockDim.x >> 5), lane);
}

bool small_fused_supports(const SmallPackDev &K, const SharedPlanDev &P) {
  static const bool on = [] { const char *e = getenv("SMPC_SMALL_FUSED"); return e && atoi(e) != 0; }();
  return on && K.M1x != nullptr && K.mp > 0 && K.mp <= NP && 2 * K.mp == P.m && P.n <= NP;
}

// same contract as launch_admm_shared_small (which classifies first when needed and calls this for paired plans)
cudaError_t launch_admm_shared_small_fused(const SmallPackDev &K, const SharedPlanDev &P, const BatchDev &Bt, const SettingsDev &S,
                                           int *queue, int *lists, int num_sms, cudaStream_t stream) {
  const int wpc = 4;
  const size_t smem = (size_t)(kFusedCtaDoubles + wpc * kFusedWarpDoubles) * sizeof(double);
  int grid = (Bt.B + wpc - 1) / wpc;
  static const int ctas_per_sm = [] { const char *e = getenv("SMPC_SMALL_CTAS"); const int v = e ? atoi(e) : kFusedCtas; return v >= 1 && v <= kFusedCtas ? v : kFusedCtas; }();
  const int resident = num_sms * ctas_per_sm;
  if (grid > resident) grid = resident;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(wpc * 32); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  static const bool quiet_on = [] { const char *e = getenv("SMPC_SMALL_QUIET"); return !e || atoi(e) != 0; }();
  const int quiet_cap = (quiet_on && lists && ctas_per_sm >= 2 && grid == resident && Bt.B <= 4 * resident * wpc) ? (num_sms + 5) / 6 : 0;
  return cudaLaunchKernelEx(&cfg, admm_shared_small_fused_kernel, K, P, Bt, S, queue, (const int *)lists, quiet_cap);
}

}  // namespace smpc
