// admm_shared_small.cu -- register-resident ADMM kernel for small QPs (n <= 16, m <= 32), the shape of
// the reference's own problem (n = 15, m = 30: mpcWindow = 15, include/ModelPredictiveControlAPI.h:26).
//
// A persistent grid of warps pulls QPs from a device-side queue.  One warp owns one QP from its first
// ADMM iteration to its last; the iterates AND the shared operators live in registers:
//   lane (h, i) = (lane >> 4, lane & 15) keeps  m1[j] = [sigma*G | W'](i, 24h + j), j < 24   (t-phase, K split in 2)
//   lane r      = lane                    keeps  wr[i] = W(r, i), i < 16                      (z-phase)
//   and its rows of xi, q̂, 1/(1+rho*lambda), z, y, l̄, ū.
// Per iteration (OSQP 0.6.x osqp_solve in plan coordinates, see admm_shared_generic.cu / plan.hpp):
//   24 DFMA + 1 shuffle-add for t, 16 DFMA for z̃, ~12 FP64 element-wise ops, two 512-byte shared-memory
//   broadcasts ([xi; w] and t).  HBM is touched once per solve (q, u in; x, y, status out), so the kernel is
//   bound by the FP64 pipe and by the latency of the dependent chain of one iteration, not by HBM.
// The element-wise tail is arranged so that only   v = alpha*z̃ + base,  z = clip(v),  w' = rho*(2z - v)
// sit on the critical path (base = (1-alpha) z + y/rho and y' = rho (v - z) are OSQP's update_z / update_y
// re-associated; y' is exactly 0 on rows that are not clipped).
// Termination checks / rho adaptation (every check_termination / adaptive_rho_interval iterations) are an
// out-of-line routine that reads its operators from shared memory.
// When the constraint rows are [G; -G] (the reference's two-sided limit, cpp:335) the PAIRED instantiation multiplies the
// top half of W only (16 + 8 DFMA per lane-iteration, lane (h, i) owns row h*mp + i; see the kernel's comment).
// Scheduling: the queue walks the difficulty classes of classify.cuh longest-expected-first; when the batch is only a few
// waves of a full grid, the hardest class runs on "quiet" SMs that keep one warp per sub-partition (see the kernel).
// Development knobs (environment, read once per process): SMPC_SMALL_QUIET=0 switches the quiet SMs off, SMPC_SMALL_CTAS=1..3
// limits the resident CTAs per SM, SMPC_SMALL_NO_PAIRS=1 (at solver creation) selects the unpaired instantiation;
// -DSMPC_SMALL_TIMELINE builds the per-instance timeline of tests/dev/dev_small_timeline.py (it overwrites obj / residuals).
#include <cstdint>
#include <cstdio>
#include <type_traits>
#include <cstdlib>

#include "classify.cuh"
#include "device_types.cuh"
#include "kernels.cuh"
#include "small_common.cuh"

namespace smpc {

namespace {

// CTA-shared operator block (doubles): VT[16][16] PVT[16][16] Ab[32][16] V[16][16] AbT[16][32]
constexpr int kCtaMatDoubles = NP * NP * 3 + 2 * MP * NP;
// per-warp block (doubles): cbuf[48] = [xi; w], tbuf[16], sbuf[32]
constexpr int kWarpDoubles = NP + MP + NP + MP;

struct CheckOut {
  double rho;              // possibly adapted
  double obj, pri_res, dua_res;
  double xbar;             // lane's entry of x̄ = V xi
  int status;              // SMPC_UNSOLVED while running
  int rho_changed;
};

__device__ __forceinline__ bool prim_inf_status(int st) { return st == SMPC_PRIMAL_INFEASIBLE || st == SMPC_PRIMAL_INFEASIBLE_INACCURATE; }
__device__ __forceinline__ bool dual_inf_status(int st) { return st == SMPC_DUAL_INFEASIBLE || st == SMPC_DUAL_INFEASIBLE_INACCURATE; }

struct LaneConst {          // per-lane constants of the plan
  double D, Dinv, E, Einv;
  int ct;
};

// OSQP update_info + check_termination (+ is_primal_infeasible / is_dual_infeasible) + adapt_rho for one QP.
// Everything it returns except xbar is uniform across the warp.  cbuf[0..15] must hold the current xi.
// PM = false: lane = row (L.E, L.Einv, z_r, ... of row `lane`).
// PM = true (row pairs [G; -G], the lane mapping of admm_shared_small_kernel<true>): lane (h, i) owns row h*mp + i for
// i < mp (L.E, L.Einv, z_r, ... of THAT row; padded lanes carry z = y = 0, l = -1, u = 1) and the products use the pairs:
// A̅'y = A̅_top'(y_top - y_bot), (A̅ v)_bot = -(A̅ v)_top, 8 terms per half-warp.
// dot of 8 operator entries op[(8h + k) * stride + i] with vec[8h + k], two accumulators (vec read as LDS.128)
__device__ __forceinline__ double half_dot8(const double *op, int stride, const double *vec, int h, int i) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(vec + 8 * h);
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const double2 v = lds128(a + 16 * j);
    s0 = fma(op[(8 * h + 2 * j) * stride + i], v.x, s0);
    s1 = fma(op[(8 * h + 2 * j + 1) * stride + i], v.y, s1);
  }
  const double s = s0 + s1;
  return s + __shfl_xor_sync(kFull, s, 16);
}
template <bool PM>
__device__ __noinline__ CheckOut check_step(const double *sm, double *cbuf, double *sbuf, const SettingsDev &S, int n, int m,
                                            double c, double cinv, LaneConst L, double rho, double qb_i, double z_r,
                                            double y_r, double dy_r, double dxi_i, double lb_r, double ub_r,
                                            bool do_check, bool approx, bool do_adapt, bool want_obj) {
  const double *sVT = sm, *sPVT = sVT + NP * NP, *sAb = sPVT + NP * NP, *sAbT = sAb + MP * NP + NP * NP;
  const int lane = threadIdx.x & 31, h = lane >> 4, i = lane & 15, r = lane;
  const bool row_ok = PM ? i < (m >> 1) : r < m;    // this lane owns a row
  const bool unscale = !S.scaled_termination;
  CheckOut o;
  o.rho = rho; o.status = SMPC_UNSOLVED; o.rho_changed = 0;
  double ax, apx, aty, Ax_r;
  if constexpr (PM) {
    const double yo = __shfl_xor_sync(kFull, y_r, 16);
    if (h == 0) sbuf[i] = y_r - yo;                  // y_top - y_bot of pair i (0 on padded lanes)
    __syncwarp();
    ax = half_dot8(sVT, NP, cbuf, h, i);
    apx = half_dot8(sPVT, NP, cbuf, h, i);
    aty = half_dot8(sAb, NP, sbuf, h, i);            // rows 0 .. mp-1 of A̅ are the top rows
    o.xbar = ax;
    __syncwarp();
    if (h == 0) sbuf[i] = ax;                        // x̄ for A̅ x̄
    __syncwarp();
    const double at = half_dot8(sAbT, MP, sbuf, h, i);
    Ax_r = row_ok ? (h ? -at : at) : 0.0;
    __syncwarp();
  } else {
    sbuf[r] = y_r;
    __syncwarp();
    ax = 0.0; apx = 0.0; aty = 0.0;
#pragma unroll
    for (int k = 0; k < NP / 2; ++k) {
      const double xk = cbuf[8 * h + k];
      ax = fma(sVT[(8 * h + k) * NP + i], xk, ax);
      apx = fma(sPVT[(8 * h + k) * NP + i], xk, apx);
    }
#pragma unroll
    for (int k = 0; k < MP / 2; ++k) aty = fma(sAb[(16 * h + k) * NP + i], sbuf[16 * h + k], aty);
    ax += __shfl_xor_sync(kFull, ax, 16);
    apx += __shfl_xor_sync(kFull, apx, 16);
    aty += __shfl_xor_sync(kFull, aty, 16);
    o.xbar = ax;
    __syncwarp();
    if (h == 0) sbuf[i] = ax;          // x̄ for A̅ x̄
    __syncwarp();
    Ax_r = 0.0;
#pragma unroll
    for (int k = 0; k < NP; ++k) Ax_r = fma(sAbT[k * MP + r], sbuf[k], Ax_r);
    __syncwarp();
  }
  const double rp = Ax_r - z_r, rd = (qb_i + apx) + aty;
  const double s_rp = wmax_bits(abits(rp)), s_z = wmax_bits(abits(z_r)), s_Ax = wmax_bits(abits(Ax_r));
  const double s_rd = wmax_bits(abits(rd)), s_q = wmax_bits(abits(qb_i)), s_Aty = wmax_bits(abits(aty)), s_Px = wmax_bits(abits(apx));
  double pri_res, dua_res, nEz, nEAx, nDq, nDAty, nDPx;
  if (unscale) {
    pri_res = wmax_bits(abits(L.Einv * rp)); nEz = wmax_bits(abits(L.Einv * z_r)); nEAx = wmax_bits(abits(L.Einv * Ax_r));
    dua_res = cinv * wmax_bits(abits(L.Dinv * rd)); nDq = wmax_bits(abits(L.Dinv * qb_i));
    nDAty = wmax_bits(abits(L.Dinv * aty)); nDPx = wmax_bits(abits(L.Dinv * apx));
  } else {
    pri_res = s_rp; nEz = s_z; nEAx = s_Ax; dua_res = s_rd; nDq = s_q; nDAty = s_Aty; nDPx = s_Px;
  }
  if (m == 0) pri_res = 0.0;
  o.obj = 0.0; o.pri_res = pri_res; o.dua_res = dua_res;

  if (do_check) {
    double ea = S.eps_abs, er = S.eps_rel, epi = S.eps_prim_inf, edi = S.eps_dual_inf;
    if (approx) { ea *= 10; er *= 10; epi *= 10; edi *= 10; }
    bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
    if (m == 0) prim_ok = true;
    else if (pri_res < ea + er * fmax(nEz, nEAx)) prim_ok = true;
    else {
      // is_primal_infeasible: project delta_y on the polar of the recession cone of [l, u]
      double d = dy_r;
      const bool uinf = ub_r > kInfty * kMinScaling, linf = lb_r < -kInfty * kMinScaling;
      if (uinf) d = linf ? 0.0 : fmin(d, 0.0); else if (linf) d = fmax(d, 0.0);
      const double nd = wmax_bits(abits(unscale ? L.E * d : d));
      if (nd > epi) {
        double lhs = 0.0;
        const double dp = fmax(d, 0.0), dm = fmin(d, 0.0);
        if (dp != 0.0) lhs += ub_r * dp;
        if (dm != 0.0) lhs += lb_r * dm;
        lhs = wsum(lhs);
        if (lhs < -epi * nd) {
          double a = 0.0;
          if constexpr (PM) {
            const double d_other = __shfl_xor_sync(kFull, d, 16);
            if (h == 0) sbuf[i] = d - d_other;
            __syncwarp();
            a = half_dot8(sAb, NP, sbuf, h, i);
          } else {
            sbuf[r] = d;
            __syncwarp();
#pragma unroll
            for (int k = 0; k < MP / 2; ++k) a = fma(sAb[(16 * h + k) * NP + i], sbuf[16 * h + k], a);
            a += __shfl_xor_sync(kFull, a, 16);
          }
          __syncwarp();
          prim_inf = wmax_bits(abits(unscale ? L.Dinv * a : a)) < epi * nd;
        }
      }
    }
    if (dua_res < ea + er * (unscale ? cinv : 1.0) * fmax(fmax(nDq, nDAty), nDPx)) dual_ok = true;
    else {
      // is_dual_infeasible on delta_x = V delta_xi
      // (the three conditions in OSQP's order, each computed only when the one before it holds: all warp-uniform)
      if (h == 0) sbuf[i] = dxi_i;
      __syncwarp();
      double dx = 0.0;
      if constexpr (PM) dx = half_dot8(sVT, NP, sbuf, h, i);
      else {
#pragma unroll
        for (int k = 0; k < NP / 2; ++k) dx = fma(sVT[(8 * h + k) * NP + i], sbuf[8 * h + k], dx);
        dx += __shfl_xor_sync(kFull, dx, 16);
      }
      const double nd = wmax_bits(abits(unscale ? L.D * dx : dx));
      const double cs = unscale ? c : 1.0;
      if (nd > edi && wsum(h == 0 ? qb_i * dx : 0.0) < -cs * edi * nd) {
        double pd = 0.0;
        if constexpr (PM) pd = half_dot8(sPVT, NP, sbuf, h, i);
        else {
#pragma unroll
          for (int k = 0; k < NP / 2; ++k) pd = fma(sPVT[(8 * h + k) * NP + i], sbuf[8 * h + k], pd);
          pd += __shfl_xor_sync(kFull, pd, 16);
        }
        if (wmax_bits(abits(unscale ? L.Dinv * pd : pd)) < cs * edi * nd) {
          __syncwarp();
          if (h == 0) sbuf[i] = dx;
          __syncwarp();
          double ad = 0.0;
          if constexpr (PM) {
            const double at = half_dot8(sAbT, MP, sbuf, h, i);
            ad = h ? -at : at;
          } else {
#pragma unroll
            for (int k = 0; k < NP; ++k) ad = fma(sAbT[k * MP + r], sbuf[k], ad);
          }
          if (unscale) ad *= L.Einv;
          const int bad = ((ub_r < kInfty * kMinScaling) && (ad > edi * nd)) || ((lb_r > -kInfty * kMinScaling) && (ad < -edi * nd));
          dual_inf = !__any_sync(kFull, bad && row_ok);
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
    const double ob = wsum(h == 0 ? 0.5 * ax * apx + qb_i * ax : 0.0);
    o.obj = prim_inf_status(o.status) ? kInfty : (dual_inf_status(o.status) ? -kInfty : (unscale ? cinv * ob : ob));
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

// Scheduling pre-pass (classify.cuh) as a stand-alone kernel: one warp per instance.
__global__ void classify_small_kernel(SmallPackDev K, SharedPlanDev P, BatchDev Bt, int *counts, int *lists) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= Bt.B) return;
  const int n = P.n, m = P.m, i = lane & 15, r = lane;
  const double q_i = (i < n && Bt.q) ? Bt.q[(size_t)b * n + i] : 0.0;
  double lo = 0.0, hi = 0.0;
  if (r < m) { lo = Bt.l ? Bt.l[(size_t)b * m + r] : P.l0[r]; hi = Bt.u ? Bt.u[(size_t)b * m + r] : P.u0[r]; }
  classify_instance(K, P, Bt.B, b, lane, q_i, lo, hi, counts, lists);
}

// PAIRED (the reference's two-sided limit [G; -G], cpp:335: row mp + i = -row i, m = 2 mp): W'w = W_top' (w_top - w_bot) and
// z̃_bot = -z̃_top, so only the top half of W is multiplied.  Lane (h, i) then owns row h*mp + i: half-warp 0 sums
// sigma*G xi and half-warp 1 W_top' wd (16 DFMA each instead of 24), both halves share the 16 products of z̃_top (8 DFMA
// each instead of 16), each lane updates its own row and the pair's difference wd = w_top - w_bot goes back to shared memory.
#ifndef SMPC_SMALL_SPLIT_Z
#define SMPC_SMALL_SPLIT_Z 1
#endif
constexpr bool kSplitZ = SMPC_SMALL_SPLIT_Z;   // PAIRED: 1 = the half-warps share z̃_top's products (8 DFMA + shuffle), 0 = both sum all 16
template <bool PAIRED>
__global__ void __launch_bounds__(128, 3)
admm_shared_small_kernel(SmallPackDev K, SharedPlanDev P, BatchDev Bt, SettingsDev S, int *queue, const int *lists, int quiet_cap) {
  constexpr int K1 = PAIRED ? NP : KH;       // terms of t summed by one half-warp
  constexpr int K2 = (PAIRED && kSplitZ) ? NP / 2 : NP;   // terms of z̃ summed by one lane
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_rank, s_ticket;
  double *sV = smem + NP * NP * 2 + MP * NP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Quiet-SM scheduling (quiet_cap > 0: the grid fills every SM with three CTAs and the batch is a few waves, so the solve
  // ends with its longest instances): the first Sq SMs to arrive keep ONE CTA (one warp per sub-partition, an
  // iteration takes ~490 instead of ~740 cycles) and those warps take the hardest class first; the other two CTAs of a
  // quiet SM leave at once.  Sq = ceil(min(size of the hardest class, 4 quiet_cap) / 4).  Only the schedule changes.
  if (quiet_cap > 0 && threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    SMPC_DBG(smid < (unsigned)kSmSlots, "smid slot");
    smid &= kSmSlots - 1;
    const int rank = atomicAdd(queue + kQRank + smid, 1);
    SMPC_DBG(rank >= 0 && rank < 3, "CTAs per SM in the quiet-SM rank counter");
    int t;
    if (rank == 0) { t = atomicAdd(queue + kQSeen, 1); atomicExch(queue + kQTicket + smid, t + 1); }
    else { do { t = atomicAdd(queue + kQTicket + smid, 0); } while (t == 0); t -= 1; }
    s_rank = rank; s_ticket = t;
  }
  double *cbuf = smem + kCtaMatDoubles + warp * kWarpDoubles, *tbuf = cbuf + NP + MP, *sbuf = tbuf + NP;
  const int h = lane >> 4, i = lane & 15;
  const int n = P.n, m = P.m, mp = m >> 1;
  // r: the row this lane owns (MP = none)
  const int r = PAIRED ? (i < mp ? h * mp + i : MP) : lane;
  const int rc = r < MP ? r : MP - 1;

  for (int e = threadIdx.x; e < NP * NP; e += blockDim.x) {
    smem[e] = K.VT[e]; smem[NP * NP + e] = K.PVT[e]; sV[e] = K.V[e];
  }
  for (int e = threadIdx.x; e < MP * NP; e += blockDim.x) {
    smem[2 * NP * NP + e] = K.Ab[e];
    const int rr = e / NP, kk = e % NP;
    smem[3 * NP * NP + MP * NP + kk * MP + rr] = K.Ab[e];   // AbT[k][r]
  }

  const double lam_i = K.lam[i];
  LaneConst LC;
  LC.D = K.D[i]; LC.Dinv = K.Dinv[i]; LC.E = K.E[rc]; LC.Einv = K.Einv[rc]; LC.ct = K.ctype[rc];   // of the lane's own row
  const double E_r = LC.E;
  const int ct_r = LC.ct;
  const double alpha = S.alpha, oma = 1.0 - S.alpha, c = P.c, cinv = P.cinv;
  const double alpha_r = (PAIRED && h) ? -alpha : alpha;   // bottom row of a pair: z̃ = -z̃_top, (-alpha) z̃_top = alpha (-z̃_top) exactly
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const uint32_t a_cv = (uint32_t)__cvta_generic_to_shared(cbuf + K1 * h);   // this half-warp's entries of [xi; w] / [xi; wd]
  const uint32_t a_tv = (uint32_t)__cvta_generic_to_shared(tbuf);
  const uint32_t a_xi = (uint32_t)__cvta_generic_to_shared(cbuf + i);
  const uint32_t a_w = (uint32_t)__cvta_generic_to_shared(cbuf + NP + (PAIRED ? i : lane));
  const uint32_t a_tz = (PAIRED && kSplitZ) ? a_tv + 8 * K2 * h : a_tv;   // this lane's terms of z̃
  const uint32_t a_t = (uint32_t)__cvta_generic_to_shared(tbuf + i);
  const int check_every = S.check_every > 0 ? S.check_every : 0x7fffffff;
  const int adapt_every = (S.adaptive_rho && S.rho_interval > 0) ? S.rho_interval : 0x7fffffff;
  __syncthreads();
  // Programmatic dependent launch: when the launcher allows it, this grid starts while the kernel before it in the stream (the
  // MPC layer's step-vector / scheduling kernel) is still running; everything above reads only the constant operator packs.
  // From here on the queue and the per-instance data written by that kernel are read: wait for it to complete (no-op otherwise).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  auto store_w = [&](double w) {
    if constexpr (PAIRED) {
      const double wo = __shfl_xor_sync(kFull, w, 16);
      if (h == 0) sts64(a_w, w - wo);
    } else sts64(a_w, w);
  };
  int n_quiet = 0;            // instances of the hardest class reserved for the quiet SMs
  bool quiet_warp = false;
  if (quiet_cap > 0 && lists != nullptr) {
    n_quiet = min(queue[1], 4 * quiet_cap);
    const int sq = (n_quiet + 3) >> 2;
    quiet_warp = s_ticket < sq && s_rank == 0;
    if (s_ticket < sq && s_rank != 0) { release_queue_warp(queue, gridDim.x * (blockDim.x >> 5), lane); return; }
  }

  for (;;) {
    // longest-expected-first: the queue walks the difficulty classes written by classify_small_kernel (the first n_quiet
    // instances of class 0 go to the quiet warps)
    int b = 0;
    if (lane == 0) {
      b = -1;
      if (quiet_warp) {
        const int qh = atomicAdd(queue + kQHard, 1);
        SMPC_DBG(qh >= 0 && n_quiet <= Bt.B, "quiet-share counter");
        if (qh < n_quiet) b = lists[qh]; else quiet_warp = false;
      }
      if (b < 0) {
        int q = atomicAdd(queue, 1);
        if (q < Bt.B - n_quiet) {
          if (lists == nullptr) b = q;
          else {
#pragma unroll
            for (int k = 0; k < kClasses; ++k) {
              const int skip = k == 0 ? n_quiet : 0, cnt = queue[1 + k] - skip;
              SMPC_DBG(cnt >= 0 && cnt + skip <= Bt.B, "class size");
              if (b < 0) { if (q < cnt) b = lists[(size_t)k * Bt.B + skip + q]; else q -= cnt; }
            }
          }
        }
      }
    }
    SMPC_DBG(b < Bt.B, "instance index from the queue");
    b = __shfl_sync(kFull, b, 0);
    if (b < 0) break;
#ifdef SMPC_SMALL_TIMELINE
    unsigned long long tl0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl0));
#endif

    // ---- load the instance (osqp_update_lin_cost / osqp_update_bounds scaling)
    const bool warm = S.warm_start && !Bt.fresh;
    const double qb_i = (i < n && Bt.q) ? c * (LC.D * Bt.q[(size_t)b * n + i]) : 0.0;
    double xi_i = (i < n && warm) ? Bt.xi[(size_t)b * n + i] : 0.0;
    double lb_r = -1.0, ub_r = 1.0, z_r = 0.0, y_r = 0.0;   // padded rows: A row = 0, never active
    if (r < m) {
      lb_r = E_r * (Bt.l ? Bt.l[(size_t)b * m + r] : P.l0[r]);
      ub_r = E_r * (Bt.u ? Bt.u[(size_t)b * m + r] : P.u0[r]);
      if (warm) { z_r = Bt.z[(size_t)b * m + r]; y_r = Bt.y[(size_t)b * m + r]; }
    }
    double rho = Bt.fresh ? fmin(fmax(S.rho0, kRhoMin), kRhoMax) : Bt.rho[b];
    int rho_updates = 0;
    // l > u (osqp_update_bounds refuses it): the instance is left UNSOLVED; a row whose class differs from the shared plan's
    // keeps the plan's rho_vec entry (see admm_shared_generic.cu)
    const bool bad_bounds = __any_sync(kFull, r < m && lb_r > ub_r);
    // q̂ = V' q̄ ; lanes of the upper half-warp start their partial sum at 0, the lower half at -q̂
    if (h == 0) sbuf[i] = qb_i;
    __syncwarp();
    double nqh_i;
    {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NP / 2; ++k) a = fma(sV[(8 * h + k) * NP + i], sbuf[8 * h + k], a);
      a += __shfl_xor_sync(kFull, a, 16);
      nqh_i = h == 0 ? -a : 0.0;
    }
    double rv = rho_row(ct_r, rho), rinv = 1.0 / rv;
    double dinv_i = 1.0 / (1.0 + rho * lam_i);
    double dxi_i = 0.0, dy_r = 0.0;
    double base_r = fma(rinv, y_r, oma * z_r), om_xi = oma * xi_i;
    __syncwarp();
    if (h == 0) sts64(a_xi, xi_i);
    store_w(rv * z_r - y_r);
    __syncwarp();

    int status = SMPC_UNSOLVED, iter = 0;
    int to_check = check_every, to_adapt = adapt_every;
    bool checked_last = false;
    CheckOut co;
    co.obj = 0.0; co.pri_res = 0.0; co.dua_res = 0.0; co.xbar = 0.0;

    if (!bad_bounds) {
      while (iter < S.max_iter) {
        // iterations until the next termination check / rho adaptation: a call-free inner loop, so the shared
        // operators stay in registers; they are (re)loaded from L1/L2 here because check_step is out of line
        int steps = to_check < to_adapt ? to_check : to_adapt;
        if (steps > S.max_iter - iter) steps = S.max_iter - iter;
        double m1[K1], wr[K2];
#pragma unroll
        for (int j = 0; j < K1; ++j) m1[j] = __ldg(K.M1T + (K1 * h + j) * NP + i);
#pragma unroll
        for (int k = 0; k < K2; ++k) {
          if constexpr (PAIRED) wr[k] = i < mp ? __ldg(K.WT + ((kSplitZ ? K2 * h : 0) + k) * MP + i) : 0.0;   // W_top(i, k)
          else wr[k] = __ldg(K.WT + k * MP + r);
        }
        // one ADMM iteration; LAST (the iteration a check follows) also leaves delta_xi, delta_y and y for check_step.
        // In between y is carried as d = v - z (y = rho d, and y / rho = d goes straight into the next base).
        double d_r = 0.0;
        bool have_d = false;
        auto iteration = [&](auto last) {
          constexpr bool LAST = decltype(last)::value;
          // ---- t = (sigma G xi + W' w - q̂) ./ (1 + rho lambda): each half-warp sums half of the concatenated terms
          double a0 = nqh_i, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
          for (int j = 0; j < K1 / 4; ++j) {
            const double2 u0 = lds128(a_cv + 32 * j), u1 = lds128(a_cv + 32 * j + 16);
            a0 = fma(m1[4 * j + 0], u0.x, a0); a1 = fma(m1[4 * j + 1], u0.y, a1);
            a2 = fma(m1[4 * j + 2], u1.x, a2); a3 = fma(m1[4 * j + 3], u1.y, a3);
          }
          double acc = (a0 + a1) + (a2 + a3);
          acc += __shfl_xor_sync(kFull, acc, 16);
          const double t_i = acc * dinv_i;
          if (h == 0) sts64(a_t, t_i);
          __syncwarp();
          // x update (off the critical path of z̃)
          const double xn = fma(alpha, t_i, om_xi);
          if constexpr (LAST) dxi_i = xn - xi_i;
          xi_i = xn; om_xi = oma * xn;
          if (h == 0) sts64(a_xi, xn);
          // ---- z̃ = W t
          double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
          for (int j = 0; j < K2 / 4; ++j) {
            const double2 u0 = lds128(a_tz + 32 * j), u1 = lds128(a_tz + 32 * j + 16);
            b0 = fma(wr[4 * j + 0], u0.x, b0); b1 = fma(wr[4 * j + 1], u0.y, b1);
            b2 = fma(wr[4 * j + 2], u1.x, b2); b3 = fma(wr[4 * j + 3], u1.y, b3);
          }
          double zt = (b0 + b1) + (b2 + b3);
          if constexpr (PAIRED && kSplitZ) zt += __shfl_xor_sync(kFull, zt, 16);   // the same bits in both halves (a + b = b + a)
          // ---- z, y updates (OSQP update_z / update_y re-associated, see the file header)
          const double v = fma(alpha_r, zt, base_r);
          const double zn = clip_sel(v, lb_r, ub_r);
          const double dn = v - zn;
          store_w(rv * fma(2.0, zn, -v));   // w' = rho z - y'
          if constexpr (LAST) {
            const double yo = have_d ? rv * d_r : y_r, yn = rv * dn;
            dy_r = yn - yo; y_r = yn;
          }
          d_r = dn; z_r = zn;
          base_r = fma(oma, zn, dn);
          __syncwarp();
        };
        for (int s = 0; s < steps - 1; ++s) { iteration(std::false_type{}); have_d = true; }
        iteration(std::true_type{});
        iter += steps; to_check -= steps; to_adapt -= steps;
        const bool do_check = to_check == 0, do_adapt = to_adapt == 0;
        checked_last = do_check;
        if (do_check) to_check = check_every;
        if (do_adapt) to_adapt = adapt_every;
        if (do_check || do_adapt) {
          co = check_step<PAIRED>(smem, cbuf, sbuf, S, n, m, c, cinv, LC, rho, qb_i, z_r, y_r, dy_r, dxi_i, lb_r, ub_r, do_check, false, do_adapt, iter >= S.max_iter);
          if (co.status != SMPC_UNSOLVED) { status = co.status; break; }
          if (co.rho_changed) {
            rho = co.rho; ++rho_updates;
            rv = rho_row(ct_r, rho); rinv = 1.0 / rv; dinv_i = 1.0 / (1.0 + rho * lam_i);
            base_r = fma(rinv, y_r, oma * z_r);
            store_w(rv * z_r - y_r);
            __syncwarp();
          }
        }
      }
      if (status == SMPC_UNSOLVED) {
        if (!checked_last) {
          co = check_step<PAIRED>(smem, cbuf, sbuf, S, n, m, c, cinv, LC, rho, qb_i, z_r, y_r, dy_r, dxi_i, lb_r, ub_r, true, false, false, true);
          status = co.status;
        }
        if (status == SMPC_UNSOLVED) {
          const CheckOut ca = check_step<PAIRED>(smem, cbuf, sbuf, S, n, m, c, cinv, LC, rho, qb_i, z_r, y_r, dy_r, dxi_i, lb_r, ub_r, true, true, false, true);
          status = ca.status == SMPC_UNSOLVED ? SMPC_MAX_ITER_REACHED : ca.status;
          if (ca.status != SMPC_UNSOLVED) co.obj = ca.obj;
        }
      }
    }
    __syncwarp();

    // ---- store_solution
    const bool has_sol = !bad_bounds && !(status == SMPC_PRIMAL_INFEASIBLE || status == SMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                                          status == SMPC_DUAL_INFEASIBLE || status == SMPC_DUAL_INFEASIBLE_INACCURATE);
    if (h == 0 && i < n) {
      if (Bt.x_out) Bt.x_out[(size_t)b * n + i] = has_sol ? LC.D * co.xbar : qnan;
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
      Bt.obj[b] = co.obj; Bt.pri_res[b] = co.pri_res; Bt.dua_res[b] = co.dua_res;
#ifdef SMPC_SMALL_TIMELINE
      // development build (tests/dev/dev_small_timeline.py): start / end time and placement of every instance
      unsigned long long tl1;
      unsigned smid_tl;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tl1));
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_tl));
      Bt.pri_res[b] = (double)(tl0 & 0xffffffffffull); Bt.dua_res[b] = (double)(tl1 & 0xffffffffffull);
      Bt.obj[b] = (double)(smid_tl * 2048 + blockIdx.x * 4 + warp) + (quiet_warp ? 0.5 : 0.0);
#endif
      if (Bt.u_apply) {
        double un = 0.0;
        if (status == SMPC_SOLVED || Bt.u_export) un = Bt.u_apply[b];
        if (status == SMPC_SOLVED) { un = __dadd_rn(un, __dmul_rn(LC.D, co.xbar)); Bt.u_apply[b] = un; }   // U += dU*[0] (cpp:105): x[0] rounded first, no FMA
        if (Bt.u_export) Bt.u_export[b] = un;                                                             // straight to the caller's (pinned) buffer
      }
      if (Bt.status_export) Bt.status_export[b] = status;
    }
  }
  // the last warp of the grid to leave re-arms the queue for the next launch (no memset node per solve)
  release_queue_warp(queue, gridDim.x * (blockDim.x >> 5), lane);
}

// =====================================================================================================================
// DMMA variant of the small-QP kernel: one CTA of four warps owns a TILE of 8 QPs ("slots").
//
// The one-warp-per-QP kernel above is bound by shared-memory wavefronts: every lane needs all 40 inputs of its two
// dot products per iteration, one broadcast wavefront per double (ncu: 58 wavefronts per QP-iteration, 75 % of the
// SM's LSU pipe at saturation).  Here the 8 slots are the N dimension of mma.sync.m8n8k4.f64, whose B operand is
// DISTRIBUTED over the lanes (one double each), so [xi; w] is read once per warp as 32 distinct doubles:
//   GEMM 1  T  = [sigma G | W'] (16 x 48) . [xi; w] (48 x 8)      warp w: row-block w & 1, K-half w >> 1, 6 DMMA,
//                                                                 partial sums through shared memory (2 KB)
//   t = (half 0 + half 1) .* dinv   (the - q̂ term is the initial accumulator of half 0)  computed by every warp
//                                                                 straight in B-fragment layout: no T panel
//   GEMM 2  Z̃  = W (32 x 16) . T (16 x 8)                          warp w: row-block w, 4 DMMA; z, y, w in registers
// with two CTA barriers and ~12 shared-memory wavefronts per QP-iteration.  The A fragments (10 doubles per lane)
// and the element-wise state of the thread's C-fragment entries stay in registers between events.
// Events (termination check / rho adaptation / max_iter, every check_termination iterations of a slot): each warp
// runs check_step -- the SAME routine as the one-warp-per-QP kernel -- for its two slots (warp, warp + 4), stores
// the QPs that finished and refills their slots from the work queue (per-problem convergence masking).
namespace {

constexpr int kSlots = 8;
// shared memory of one CTA (doubles): operator block of check_step | per-warp cbuf[16] + sbuf[32] | panels [row][8]
constexpr int kMmaWarpDoubles = NP + MP;
constexpr int oS = 0;                        // [xi (16 rows); w (32 rows)]
constexpr int oP0 = oS + (NP + MP) * 8;      // GEMM 1 partial sums, K-half 0 / 1
constexpr int oP1 = oP0 + NP * 8;
constexpr int oDx = oP1 + NP * 8;            // delta_xi of the event iteration
constexpr int oQh = oDx + NP * 8;            // q̂ = V' q̄
constexpr int oQb = oQh + NP * 8;            // q̄ = c D q
constexpr int oDv = oQb + NP * 8;            // 1 / (1 + rho lambda)
constexpr int oZ = oDv + NP * 8;
constexpr int oY = oZ + MP * 8;
constexpr int oLb = oY + MP * 8;
constexpr int oUb = oLb + MP * 8;
constexpr int kMmaPanelDoubles = oUb + MP * 8;

struct MmaCtl {
  double rho[kSlots];
  int inst[kSlots];      // QP index of the slot, -1 = empty
  int it0[kSlots];       // tile iteration at which the slot's QP started
  int rho_up[kSlots];
  int next_event, active;
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
      : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// next QP of the work queue (index order, or longest-expected-first through the class lists); -1 when it is empty
__device__ __forceinline__ int pop_instance(int *queue, const int *lists, int B) {
  int q = atomicAdd(queue, 1);
  if (q >= B) return -1;
  if (lists == nullptr) return q;
#pragma unroll
  for (int k = 0; k < kClasses; ++k) {
    const int cnt = queue[1 + k];
    if (q < cnt) return lists[(size_t)k * B + q];
    q -= cnt;
  }
  return -1;
}

}  // namespace

namespace {
template <int OFF> __device__ __forceinline__ double lds64o(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
template <int OFF> __device__ __forceinline__ double2 lds128o(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(addr), "n"(OFF) : "memory");
  return v;
}
template <int OFF> __device__ __forceinline__ void sts64o(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(v) : "memory");
}
template <int OFF> __device__ __forceinline__ void sts128o(uint32_t addr, double a, double b) {
  asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(addr), "n"(OFF), "d"(a), "d"(b) : "memory");
}
}  // namespace

__global__ void __launch_bounds__(128, 4)
admm_shared_small_mma_kernel(SmallPackDev K, SharedPlanDev P, BatchDev Bt, SettingsDev S, int *queue, const int *lists) {
  extern __shared__ __align__(16) double smem[];
  double *sV = smem + NP * NP * 2 + MP * NP;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *cbuf = smem + kCtaMatDoubles + warp * kMmaWarpDoubles, *sbuf = cbuf + NP;
  double *pan = smem + kCtaMatDoubles + 4 * kMmaWarpDoubles;
  MmaCtl &C = *reinterpret_cast<MmaCtl *>(pan + kMmaPanelDoubles);
  const int h = lane >> 4, i = lane & 15, r = lane;          // check_step roles (one slot per pass)
  const int g = lane >> 2, q2 = 2 * (lane & 3);               // DMMA roles: C fragment = rows 8 rb + g, slots q2, q2 + 1
  const int bfrag = (lane & 3) * 8 + g;                       // B fragment element: k-row lane & 3, slot g
  const int n = P.n, m = P.m;

  for (int e = tid; e < NP * NP; e += blockDim.x) {
    smem[e] = K.VT[e]; smem[NP * NP + e] = K.PVT[e]; sV[e] = K.V[e];
  }
  for (int e = tid; e < MP * NP; e += blockDim.x) {
    smem[2 * NP * NP + e] = K.Ab[e];
    const int rr = e / NP, kk = e % NP;
    smem[3 * NP * NP + MP * NP + kk * MP + rr] = K.Ab[e];   // AbT[k][r]
  }
  for (int e = tid; e < kMmaPanelDoubles; e += blockDim.x) pan[e] = 0.0;
  if (tid < kSlots) { C.inst[tid] = -1; C.it0[tid] = 0; C.rho_up[tid] = 0; C.rho[tid] = 1.0; }
  if (tid == 0) { C.next_event = 0; C.active = 0; }

  const double alpha = S.alpha, oma = 1.0 - S.alpha;
  const int check_every = S.check_every > 0 ? S.check_every : 0, adapt_every = (S.adaptive_rho && S.rho_interval > 0) ? S.rho_interval : 0;
  const bool warm = S.warm_start && !Bt.fresh;

  // GEMM 1 unit of this warp = row-block warp & 1, k-steps 6 (warp >> 1) .. +5 of [sigma G | W'] (K = 48); GEMM 2 unit =
  // row-block warp of W (K = 16).  The A fragments (a1, a2) are (re)loaded after every event so that they are not live
  // across the event code.
  const int rb1 = warp & 1, kh = warp >> 1;
  double a1[6], a2[4];
  const int ct_c = K.ctype[8 * warp + g];                     // row class of this thread's C-fragment row

  // loop-invariant shared-memory addresses (32-bit shared window, byte offsets as immediates)
  const uint32_t sPan = (uint32_t)__cvta_generic_to_shared(pan);
  const uint32_t aB = sPan + 8 * bfrag;                                      // B-fragment element of a panel row group
  const uint32_t aB1 = aB + 8 * 192 * kh;                                    // GEMM 1 B fragments of this K-half: + 256 j
  const uint32_t aPartSt = sPan + 8 * ((kh ? oP1 : oP0) + (8 * rb1 + g) * 8 + q2);
  const uint32_t aXi = aB + 8 * 32 * warp;                                   // this warp updates xi rows 4 warp .. 4 warp + 3
  const uint32_t aC = sPan + 8 * ((8 * warp + g) * 8 + q2);                  // C-fragment offset inside an m-panel
  const bool sel1 = warp & 1, sel2 = warp & 2;

  // element-wise state in registers between events
  double r_nq[2], r_dv[4], r_xi = 0.0, r_omxi = 0.0, r_y[2], r_lo[2], r_hi[2], r_rv[2], r_ri[2], r_base[2];
  int k = 0, next_event = 0;
  __syncthreads();

#ifdef SMPC_PROFILE
  long long pf_ev = 0, pf_it = 0, pf_t0 = clock64(), pf_chk = 0; int pf_nev = 0;
#endif
  for (;;) {
    if (k == next_event) {
#ifdef SMPC_PROFILE
      { const long long tt = clock64(); pf_it += tt - pf_t0; pf_t0 = tt; }
#endif
      // ================= event: every warp handles the slots warp and warp + 4
      // (per-lane plan constants are loaded here, not kept in registers across the main loop)
      const double lam_i = __ldg(K.lam + i), c = P.c, cinv = P.cinv;
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
      LaneConst LC;
      LC.D = __ldg(K.D + i); LC.Dinv = __ldg(K.Dinv + i); LC.E = __ldg(K.E + r); LC.Einv = __ldg(K.Einv + r); LC.ct = __ldg(K.ctype + r);
      for (int s = warp; s < kSlots; s += 4) {
        int b = C.inst[s];
        double rho = C.rho[s];
        bool refill = b < 0 && k == 0;            // the first event only fills the tile
        if (b >= 0) {
          const int li = k - C.it0[s];
          const bool do_check = check_every && li % check_every == 0, do_adapt = adapt_every && li % adapt_every == 0;
          const bool at_max = li >= S.max_iter;
          if (do_check || do_adapt || at_max) {
            // this slot's column of the panels in the one-warp-per-QP register layout of check_step
            const double qb_i = pan[oQb + i * 8 + s], dxi_i = pan[oDx + i * 8 + s], xi_i = pan[oS + i * 8 + s];
            const double z_r = pan[oZ + r * 8 + s], y_r = pan[oY + r * 8 + s], dy_r = pan[oS + (NP + r) * 8 + s];
            const double lb_r = pan[oLb + r * 8 + s], ub_r = pan[oUb + r * 8 + s];
            if (h == 0) cbuf[i] = xi_i;
            __syncwarp();
            int status = SMPC_UNSOLVED;
            CheckOut co;
            co.obj = 0.0; co.pri_res = 0.0; co.dua_res = 0.0; co.xbar = 0.0;
            if (do_check || do_adapt) {
#ifdef SMPC_PROFILE
              const long long tc = clock64();
#endif
              co = check_step<false>(smem, cbuf, sbuf, S, n, m, c, cinv, LC, rho, qb_i, z_r, y_r, dy_r, dxi_i, lb_r, ub_r, do_check, false, do_adapt, at_max);
              status = co.status;
#ifdef SMPC_PROFILE
              pf_chk += clock64() - tc;
#endif
              if (status == SMPC_UNSOLVED && co.rho_changed) {
                rho = co.rho;
                if (lane == 0) { C.rho[s] = rho; C.rho_up[s]++; }
                if (h == 0) pan[oDv + i * 8 + s] = 1.0 / (1.0 + rho * lam_i);
              }
            }
            if (at_max && status == SMPC_UNSOLVED) {
              if (!do_check) {
                co = check_step<false>(smem, cbuf, sbuf, S, n, m, c, cinv, LC, rho, qb_i, z_r, y_r, dy_r, dxi_i, lb_r, ub_r, true, false, false, true);
                status = co.status;
              }
              if (status == SMPC_UNSOLVED) {
                const CheckOut ca = check_step<false>(smem, cbuf, sbuf, S, n, m, c, cinv, LC, rho, qb_i, z_r, y_r, dy_r, dxi_i, lb_r, ub_r, true, true, false, true);
                status = ca.status == SMPC_UNSOLVED ? SMPC_MAX_ITER_REACHED : ca.status;
                if (ca.status != SMPC_UNSOLVED) co.obj = ca.obj;
              }
            }
            __syncwarp();
            if (status != SMPC_UNSOLVED) {
              // ---- store_solution
              const bool has_sol = !(status == SMPC_PRIMAL_INFEASIBLE || status == SMPC_PRIMAL_INFEASIBLE_INACCURATE ||
                                     status == SMPC_DUAL_INFEASIBLE || status == SMPC_DUAL_INFEASIBLE_INACCURATE);
              if (h == 0 && i < n) {
                if (Bt.x_out) Bt.x_out[(size_t)b * n + i] = has_sol ? LC.D * co.xbar : qnan;
                Bt.xi[(size_t)b * n + i] = has_sol ? xi_i : 0.0;
              }
              if (r < m) {
                if (Bt.y_out) Bt.y_out[(size_t)b * m + r] = has_sol ? cinv * (LC.E * y_r) : qnan;
                Bt.z[(size_t)b * m + r] = has_sol ? z_r : 0.0;
                Bt.y[(size_t)b * m + r] = has_sol ? y_r : 0.0;
              }
              if (lane == 0) {
                Bt.rho[b] = rho;
                Bt.status[b] = status; Bt.iter[b] = li; Bt.rho_updates[b] = C.rho_up[s];
                Bt.obj[b] = co.obj; Bt.pri_res[b] = co.pri_res; Bt.dua_res[b] = co.dua_res;
                if (Bt.u_apply && status == SMPC_SOLVED) Bt.u_apply[b] = __dadd_rn(Bt.u_apply[b], __dmul_rn(LC.D, co.xbar));   // U += dU*[0] (cpp:105)
              }
              refill = true;
            }
          }
        }
        if (refill) {
          // ---- next QP of the queue into slot s (osqp_update_lin_cost / osqp_update_bounds scaling); QPs with invalid
          //      bounds (see admm_shared_generic.cu) are stored as UNSOLVED at once and the slot is filled again
          for (;;) {
            b = 0;
            if (lane == 0) b = pop_instance(queue, lists, Bt.B);
            b = __shfl_sync(kFull, b, 0);
            double qb_i = 0.0, xi_i = 0.0, lb_r = -1.0, ub_r = 1.0, z_r = 0.0, y_r = 0.0, nqh = 0.0;
            rho = 1.0;
            if (b >= 0) {
              qb_i = (i < n && Bt.q) ? c * (LC.D * Bt.q[(size_t)b * n + i]) : 0.0;
              if (i < n && warm) xi_i = Bt.xi[(size_t)b * n + i];
              if (r < m) {
                lb_r = LC.E * (Bt.l ? Bt.l[(size_t)b * m + r] : P.l0[r]);
                ub_r = LC.E * (Bt.u ? Bt.u[(size_t)b * m + r] : P.u0[r]);
                if (warm) { z_r = Bt.z[(size_t)b * m + r]; y_r = Bt.y[(size_t)b * m + r]; }
              }
              rho = Bt.fresh ? fmin(fmax(S.rho0, kRhoMin), kRhoMax) : Bt.rho[b];
              if (__any_sync(kFull, r < m && lb_r > ub_r)) {
                if (h == 0 && i < n) { if (Bt.x_out) Bt.x_out[(size_t)b * n + i] = qnan; Bt.xi[(size_t)b * n + i] = 0.0; }
                if (r < m) { if (Bt.y_out) Bt.y_out[(size_t)b * m + r] = qnan; Bt.z[(size_t)b * m + r] = 0.0; Bt.y[(size_t)b * m + r] = 0.0; }
                if (lane == 0) {
                  Bt.rho[b] = rho; Bt.status[b] = SMPC_UNSOLVED; Bt.iter[b] = 0; Bt.rho_updates[b] = 0;
                  Bt.obj[b] = 0.0; Bt.pri_res[b] = 0.0; Bt.dua_res[b] = 0.0;
                }
                continue;
              }
              // q̂ = V' q̄
              if (h == 0) sbuf[i] = qb_i;
              __syncwarp();
              double a = 0.0;
#pragma unroll
              for (int kk = 0; kk < NP / 2; ++kk) a = fma(sV[(8 * h + kk) * NP + i], sbuf[8 * h + kk], a);
              a += __shfl_xor_sync(kFull, a, 16);
              nqh = a;
              __syncwarp();
            }
            if (h == 0) {
              pan[oS + i * 8 + s] = xi_i; pan[oQb + i * 8 + s] = qb_i; pan[oQh + i * 8 + s] = nqh; pan[oDx + i * 8 + s] = 0.0;
              pan[oDv + i * 8 + s] = b >= 0 ? 1.0 / (1.0 + rho * lam_i) : 1.0;
            }
            pan[oZ + r * 8 + s] = z_r; pan[oY + r * 8 + s] = y_r; pan[oLb + r * 8 + s] = lb_r; pan[oUb + r * 8 + s] = ub_r;
            if (lane == 0) { C.inst[s] = b; C.it0[s] = k; C.rho_up[s] = 0; C.rho[s] = rho; }
            break;
          }
        }
        // ---- w = rho_vec z - y (the w panel carried delta_y during the event iteration)
        __syncwarp();
        pan[oS + (NP + r) * 8 + s] = rho_row(LC.ct, rho) * pan[oZ + r * 8 + s] - pan[oY + r * 8 + s];
      }
      __syncthreads();
      if (tid == 0) {
        int act = 0, ne = 0x7fffffff;
        for (int s = 0; s < kSlots; ++s) {
          if (C.inst[s] < 0) continue;
          ++act;
          const int li = k - C.it0[s];
          int nx = S.max_iter;
          if (check_every) nx = min(nx, (li / check_every + 1) * check_every);
          if (adapt_every) nx = min(nx, (li / adapt_every + 1) * adapt_every);
          ne = min(ne, C.it0[s] + nx);
        }
        C.active = act; C.next_event = ne;
      }
      __syncthreads();
#ifdef SMPC_PROFILE
      { const long long tt = clock64(); pf_ev += tt - pf_t0; pf_t0 = tt; ++pf_nev; }
      if (C.active == 0 && blockIdx.x == 0 && tid == 0) printf("mma profile: %d events %lld cycles (check_step %lld), %d iterations %lld cycles\n", pf_nev, pf_ev, pf_chk, k, pf_it);
#endif
      if (C.active == 0) break;
      next_event = C.next_event;
      // ---- reload this thread's registers from the panels; operator fragments from L1 / L2 (M1T, WT are k-major:
      //      fragment element (row 8 rb + g, k = 4 ks + (lane & 3)))
#pragma unroll
      for (int j = 0; j < 6; ++j) a1[j] = __ldg(K.M1T + (4 * (6 * kh + j) + (lane & 3)) * NP + 8 * rb1 + g);
#pragma unroll
      for (int j = 0; j < 4; ++j) a2[j] = __ldg(K.WT + (4 * j + (lane & 3)) * MP + 8 * warp + g);
      r_dv[0] = lds64o<8 * oDv>(aB); r_dv[1] = lds64o<8 * (oDv + 32)>(aB); r_dv[2] = lds64o<8 * (oDv + 64)>(aB); r_dv[3] = lds64o<8 * (oDv + 96)>(aB);
      r_xi = lds64o<8 * oS>(aXi); r_omxi = oma * r_xi;
      {
        const double2 qv = *reinterpret_cast<const double2 *>(pan + oQh + (8 * rb1 + g) * 8 + q2);
        r_nq[0] = kh ? 0.0 : -qv.x; r_nq[1] = kh ? 0.0 : -qv.y;
        const double2 zz = lds128o<8 * oZ>(aC), yy = lds128o<8 * oY>(aC), lo = lds128o<8 * oLb>(aC), hi = lds128o<8 * oUb>(aC);
        const double r_z[2] = {zz.x, zz.y};
        r_y[0] = yy.x; r_y[1] = yy.y; r_lo[0] = lo.x; r_lo[1] = lo.y; r_hi[0] = hi.x; r_hi[1] = hi.y;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          r_rv[j] = rho_row(ct_c, C.rho[q2 + j]); r_ri[j] = 1.0 / r_rv[j];
          r_base[j] = fma(r_ri[j], r_y[j], oma * r_z[j]);
        }
      }
    }

    // ================= one ADMM iteration of the tile (same arithmetic per QP as admm_shared_small_kernel)
    ++k;
    const bool event = k == next_event;
    // ---- GEMM 1 partial sums
    {
      const double b0 = lds64o<0>(aB1), b1 = lds64o<256>(aB1), b2 = lds64o<512>(aB1);
      const double b3 = lds64o<768>(aB1), b4 = lds64o<1024>(aB1), b5 = lds64o<1280>(aB1);
      double c0[2] = {r_nq[0], r_nq[1]}, c1[2] = {0.0, 0.0};
      dmma(c0, a1[0], b0); dmma(c1, a1[1], b1); dmma(c0, a1[2], b2);
      dmma(c1, a1[3], b3); dmma(c0, a1[4], b4); dmma(c1, a1[5], b5);
      sts128o<0>(aPartSt, c0[0] + c1[0], c0[1] + c1[1]);
    }
    __syncthreads();
    // ---- t in B-fragment layout, x update by the warp that owns the rows
    double t[4];
    t[0] = (lds64o<8 * oP0>(aB) + lds64o<8 * oP1>(aB)) * r_dv[0];
    t[1] = (lds64o<8 * (oP0 + 32)>(aB) + lds64o<8 * (oP1 + 32)>(aB)) * r_dv[1];
    t[2] = (lds64o<8 * (oP0 + 64)>(aB) + lds64o<8 * (oP1 + 64)>(aB)) * r_dv[2];
    t[3] = (lds64o<8 * (oP0 + 96)>(aB) + lds64o<8 * (oP1 + 96)>(aB)) * r_dv[3];
    // ---- GEMM 2 and the z, y, w updates (OSQP update_z / update_y re-associated as in the one-warp kernel)
    {
      double c0[2] = {0.0, 0.0}, c1[2] = {0.0, 0.0};
      dmma(c0, a2[0], t[0]); dmma(c1, a2[1], t[1]); dmma(c0, a2[2], t[2]); dmma(c1, a2[3], t[3]);
      {
        const double ta = sel1 ? t[1] : t[0], tb = sel1 ? t[3] : t[2], tw = sel2 ? tb : ta;
        const double xn = fma(alpha, tw, r_omxi);
        sts64o<8 * oS>(aXi, xn);
        if (event) sts64o<8 * oDx>(aXi, xn - r_xi);
        r_xi = xn; r_omxi = oma * xn;
      }
      double wv[2], r_z[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const double v = fma(alpha, c0[j] + c1[j], r_base[j]);
        const double zn = v < r_lo[j] ? r_lo[j] : (v > r_hi[j] ? r_hi[j] : v);
        const double yn = r_rv[j] * (v - zn);
        wv[j] = event ? yn - r_y[j] : r_rv[j] * fma(2.0, zn, -v);   // before an event the w panel carries delta_y
        r_y[j] = yn; r_z[j] = zn;
        r_base[j] = fma(r_ri[j], yn, oma * zn);
      }
      sts128o<8 * (oS + NP * 8)>(aC, wv[0], wv[1]);
      if (event) { sts128o<8 * oZ>(aC, r_z[0], r_z[1]); sts128o<8 * oY>(aC, r_y[0], r_y[1]); }
    }
    __syncthreads();
  }
  if (tid == 0) release_queue(queue, gridDim.x);
}

bool small_kernel_supports(int n, int m) { return n >= 1 && n <= NP && m >= 0 && m <= MP; }

size_t small_pack_doubles() { return (size_t)(NP + MP) * NP + NP * MP + 3 * NP * NP + MP * NP + 3 * NP + 2 * MP; }

int small_queue_ints() { return kQueueInts; }
int small_sched_classes() { return kClasses; }

// scheduling pre-pass: fills the class lists and sizes (the queue is zero on entry: cleared at upload and re-armed by
// the last warp of every solve, see release_queue)
static cudaError_t launch_classify_small(const SmallPackDev &K, const SharedPlanDev &P, const BatchDev &Bt, int *queue, int *lists,
                                         cudaStream_t stream) {
  classify_small_kernel<<<(Bt.B + 7) / 8, 256, 0, stream>>>(K, P, Bt, queue + 1, lists);
  return cudaGetLastError();
}

// DMMA variant: 8 QPs per CTA of four warps, up to four CTAs per SM
cudaError_t launch_admm_shared_small_mma(const SmallPackDev &K, const SharedPlanDev &P, const BatchDev &Bt,
                                         const SettingsDev &S, int *queue, int *lists, bool classified, int num_sms, cudaStream_t stream) {
  const size_t smem = (size_t)(kCtaMatDoubles + 4 * kMmaWarpDoubles + kMmaPanelDoubles) * sizeof(double) + sizeof(MmaCtl);
  cudaError_t e = (lists && !classified) ? launch_classify_small(K, P, Bt, queue, lists, stream) : cudaSuccess;
  if (e != cudaSuccess) return e;
  int grid = (Bt.B + kSlots - 1) / kSlots;
  if (grid > num_sms * 4) grid = num_sms * 4;
  // (the attribute is per device: set it on every launch rather than once per process)
  e = cudaFuncSetAttribute(admm_shared_small_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  admm_shared_small_mma_kernel<<<grid, 128, smem, stream>>>(K, P, Bt, S, queue, lists);
  return cudaGetLastError();
}

// queue: [0] work counter, [1..kClasses] class sizes; lists: kClasses * B instance indices (nullptr = index order)
cudaError_t launch_admm_shared_small(const SmallPackDev &K, const SharedPlanDev &P, const BatchDev &Bt,
                                     const SettingsDev &S, int *queue, int *lists, bool classified, int num_sms, cudaStream_t stream) {
  const int wpc = 4;
  const size_t smem = (size_t)(kCtaMatDoubles + wpc * kWarpDoubles) * sizeof(double);
  cudaError_t e = (lists && !classified) ? launch_classify_small(K, P, Bt, queue, lists, stream) : cudaSuccess;
  if (e != cudaSuccess) return e;
  // paired rows [G; -G] with at most 16 pairs: the one-phase kernel (admm_shared_small_fused.cu)
  if (small_fused_supports(K, P)) return launch_admm_shared_small_fused(K, P, Bt, S, queue, lists, num_sms, stream);
  int grid = (Bt.B + wpc - 1) / wpc;
  static const int ctas_per_sm = [] { const char *e = getenv("SMPC_SMALL_CTAS"); const int v = e ? atoi(e) : 3; return v >= 1 && v <= 3 ? v : 3; }();
  const int resident = num_sms * ctas_per_sm;   // __launch_bounds__(128, 3): at most three CTAs (12 warps) per SM
  if (grid > resident) grid = resident;
  // programmatic stream serialization: the prologue overlaps the tail of the preceding kernel (which must trigger it with
  // griddepcontrol.launch_dependents, as the MPC layer's kernels do; otherwise this is an ordinary serialized launch)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(wpc * 32); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  // quiet SMs for the hardest class when the batch is a few waves of a full grid (the tail of the longest instances sets
  // the time); deeper batches are throughput-bound and keep every warp slot busy
  static const bool quiet_on = [] { const char *e = getenv("SMPC_SMALL_QUIET"); return !e || atoi(e) != 0; }();
  const int quiet_cap = (quiet_on && lists && ctas_per_sm == 3 && grid == resident && Bt.B <= 4 * resident * wpc) ? (num_sms + 5) / 6 : 0;
  if (K.mp > 0 && 2 * K.mp == P.m) return cudaLaunchKernelEx(&cfg, admm_shared_small_kernel<true>, K, P, Bt, S, queue, (const int *)lists, quiet_cap);
  return cudaLaunchKernelEx(&cfg, admm_shared_small_kernel<false>, K, P, Bt, S, queue, (const int *)lists, quiet_cap);
}

}  // namespace smpc
