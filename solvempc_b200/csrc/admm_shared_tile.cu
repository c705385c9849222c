// admm_shared_tile.cu -- shared-factor ADMM kernel for mid-size QPs (n > 16: horizons 30..100, the 12-state
// quadrotor), the "(n x n) . (n x batch) contraction" of the north star, on the FP64 tensor pipe (DMMA).
//
// One CTA owns a TILE of TB = 8*NB QPs ("slots") and keeps every iterate of the tile in shared memory for the
// whole solve; HBM is touched once per solve (q, l, u in; x, y, status out).  The shared operators of the plan
// (plan.hpp: sigma*G, W = A̅V, ...) are streamed from L2 once per tile-iteration as pre-packed
// mma.sync.m8n8k4.f64 A-fragments (512 contiguous bytes per warp load) and multiplied against the tile's iterate
// PANELS (the 8 slots of a panel are the N dimension of the DMMA).  Per iteration (OSQP 0.6.x osqp_solve in
// plan coordinates, reference call site src/ModelPredictiveControlAPI.cpp:102; same maths as admm_shared_generic.cu):
//     T  = ([sigma*G | W'] [xi; w] - q̂) .* dinv            GEMM 1, K = n + m      dinv = 1/(1 + rho_b*lambda_i)
//     xi'= alpha T + (1-alpha) xi                           (after the barrier, by the warp that owns the rows)
//     Z̃  = W T                                              GEMM 2, K = n
//     z, y, w = rho_vec z - y                               epilogue of GEMM 2 (clip to [l̄, ū], dual update)
// with two CTA barriers per iteration.  Slots are independent QPs: each has its own rho, iteration counter and
// status; termination checks / rho adaptation (four more panel GEMMs: x̄ = V xi, P̄x, A̅'y, A̅x̄) run when a slot is due.
// A slot whose QP finished is stored and REFILLED from a device-wide work queue at that check, so a tile never
// idles on its slowest member (per-problem convergence masking = slot recycling).
//
// Panel layout in shared memory: [nb][row][8] doubles (8-slot blocks), so a DMMA B-fragment load (4 k-rows x 8 slots)
// is 32 consecutive doubles and the C-fragment of a thread is a double2 of the same panel layout.  Inside a row the slots are
// SWIZZLED: slot s of row r sits at position s ^ (4 * ((r >> 1) & 1)).  A 64-bit shared-memory load serves 16 lanes per
// wavefront; in a B fragment those are k-rows 0..3 x slots 0..3, and unswizzled rows 0 / 2 (and 1 / 3) are 128 bytes apart =
// the same banks: every B-fragment load took 4 wavefronts instead of 2 (37 % of all shared-memory wavefronts of the kernel,
// ncu).  With the swizzle the 16 lanes touch 16 distinct 8-byte bank pairs; pairs of slots (2q, 2q + 1) stay adjacent.
#include <cstdint>
#include <cstdlib>
#include <cstdio>

#include "device_types.cuh"
#include "kernels.cuh"

namespace smpc {

namespace {

constexpr int kTileWarps = 8;
constexpr int kTileThreads = kTileWarps * 32;
constexpr int kRG = 4;          // row-blocks (8 operator rows each) a warp accumulates at once
// A-fragment prefetch distance in k-pairs.  Measured (round 2, tools/ab_tile.sh + tools/ab_run.sh, B200; config 3 ms / config 5 s per 100
// closed-loop steps): 1 -> 96.7 / 0.671, 2 -> 94.2 / 0.686, 3 -> 95.4 / 0.683, 4 -> 98.6 / 0.705, 5 -> 104.9 / 0.731, 6 -> 105.9 / 0.785,
// 8 (spills) -> 127.3 / 0.892: the GEMMs do not wait for the operator stream's latency.  Diagnostic builds that replace the operator loads
// (SMPC_TILE_DIAG_NOLDG), the B-fragment loads (SMPC_TILE_DIAG_NOLDS) or both by register values run the n = 200 iteration in 15.7 / 17.8 /
// 15.8 k cycles against 18.0 k: the whole operator stream is worth 13 %, the rest of the GEMM phases is DMMA issue
// (profiles/microbench/tile_gemm_shape.cu: the bare DMMA + barrier skeleton of one n = 200 GEMM takes 6.1 k cycles, 17.4 cycles per DMMA on the
// busiest sub-partition against the pipe's 16; the kernel's GEMM + epilogue 8.1-8.8 k, 7.1 k without operator loads).  Also measured and
// not kept: the next GEMM's first fragments loaded before the epilogue and the barrier (+3 % / +4 %), B fragments one k-pair ahead
// (+-0), separate accumulators for the two k-steps of a pair (+2 % / +3 %), every tile prefetching its own share of the ticket queue
// into L2 at an event (+-0 / +5 %), deeper unrolling of the refill loads and stores (+-0).
#ifndef SMPC_TILE_UNROLL_LD
#define SMPC_TILE_UNROLL_LD 4
#endif
#ifndef SMPC_TILE_UNROLL_ST
#define SMPC_TILE_UNROLL_ST 4
#endif
// (per kernel variant, below: kRing; SMPC_TILE_RING overrides it for every variant)
constexpr int kUnrollLd = SMPC_TILE_UNROLL_LD, kUnrollSt = SMPC_TILE_UNROLL_ST;

enum NormId {
  N_RP_S, N_Z_S, N_AX_S, N_RP_U, N_Z_U, N_AX_U,
  N_RD_S, N_Q_S, N_ATY_S, N_PX_S, N_RD_U, N_Q_U, N_ATY_U, N_PX_U,
  N_DY, N_ATD, N_DX, N_PD, N_DXI, N_COUNT
};
enum SumId { S_OBJ, S_LHS, S_QD, S_COUNT };

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 ldg_stream(const double2 *p) {
  double2 v;
#ifdef SMPC_TILE_DIAG_NOLDG   // diagnostic only (wrong results): no operator stream, what do the GEMMs cost without it?
  return make_double2(__longlong_as_double(0x3F50000000000000LL | (long long)((size_t)p & 0xFF0)), 1e-3);   // (integer ops only: no conversions in the loop)
#endif
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int swz(int row) { return ((row >> 1) & 1) << 2; }   // panel slot swizzle (see the file header)
__device__ __forceinline__ double rho_of(int ct, double rho) { return ct == 0 ? rho : (ct == 1 ? kRhoEqOverIneq * rho : kRhoMin); }

// per-tile bookkeeping in shared memory
template <int TB, int W = kTileWarps>
struct TileCtl {
  unsigned long long nmax[N_COUNT][TB];    // max-norms as IEEE bit patterns of non-negative doubles
  double psum[S_COUNT][W][TB];
  double rho[TB], rinv[TB], rho_eq[TB], rinv_eq[TB];
  double obj[TB], pri[TB], dua[TB];
  int inst[TB];        // QP index of the slot, -1 = empty
  int it0[TB];         // tile iteration at which the slot's QP started
  int rho_up[TB];
  int status[TB];
  int flags[TB];       // scratch bit flags per slot (F_*)
  int next_event;      // next tile iteration at which some slot is due for a check / adaptation / max_iter
  int active;          // number of occupied slots
  int refill_mask, work_flag;
};
enum SlotFlag { F_DUE_CHECK = 1, F_DUE_ADAPT = 2, F_AT_MAX = 4, F_PRIM_OK = 8, F_DUAL_OK = 16, F_NEED_PINF = 32,
                F_NEED_DINF = 64, F_BADROW = 128, F_DONE = 256, F_NEED_ATD = 512, F_NEED_DINF2 = 1024, F_RHO_NEW = 2048,
                F_BADBOUNDS = 4096, F_CADENCE = 8192 };

// acc[r][nb] += Op[row-blocks rb .. rb+NR) (k-pairs kp0 .. kp0+cnt) * panel.  opl = pack + lane; kpt = k-pairs per
// operator row-block; bp = panel + (lane&3)*8 + (lane>>2) (+ 64*first k-pair); nbs = doubles between 8-slot blocks.
// The A fragments run kRing k-pairs ahead of their use in a register ring (covers the L2 latency); the main loop is
// branch-free (prefetches past the end are clamped to the last k-pair).
// PB ("paired B"): the B operand is the difference of two panel rows, bp[...] - bp2[...] (A̅'y = G'(y_top - y_bot) for row pairs [G; -G]).
template <int NB, int NR, int RG, int kRing, bool PB = false>
__device__ __forceinline__ void gemm_run(const double2 *__restrict__ opl, int kpt, int rb, int kp0, int cnt,
                                         const double *bp, int nbs, double (&acc)[RG][NB][2], const double *bp2 = nullptr, int rowlim = 0) {
  const double2 *ap[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) ap[r] = opl + ((size_t)(rb + r) * kpt + kp0) * 32;
  double2 ring[kRing][NR];
#pragma unroll
  for (int d = 0; d < kRing; ++d) {
    const int kk = min(d, cnt - 1) * 32;
#pragma unroll
    for (int r = 0; r < NR; ++r) ring[d][r] = ldg_stream(ap[r] + kk);
  }
  int kp = 0;
  for (; kp + kRing <= cnt; kp += kRing) {
#pragma unroll
    for (int d = 0; d < kRing; ++d) {
      double b[NB][2];
#ifdef SMPC_TILE_DIAG_NOLDS   // diagnostic only (wrong results): no B-fragment loads
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) { b[nb][0] = __longlong_as_double(0x3F50000000000000LL | ((long long)(kp + d + nb) << 8)); b[nb][1] = 1e-3; }
#else
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        b[nb][0] = bp[nb * nbs + d * 64]; b[nb][1] = bp[nb * nbs + d * 64 + 32];
        if constexpr (PB) {   // (rows past the last pair are padding of the operator: keep them finite whatever lies behind the panel)
          b[nb][0] = 8 * (kp + d) < rowlim ? b[nb][0] - bp2[nb * nbs + d * 64] : 0.0;
          b[nb][1] = 8 * (kp + d) + 4 < rowlim ? b[nb][1] - bp2[nb * nbs + d * 64 + 32] : 0.0;
        }
      }
#endif
      double2 a[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) a[r] = ring[d][r];
      const int kk = min(kp + d + kRing, cnt - 1) * 32;
#pragma unroll
      for (int r = 0; r < NR; ++r) ring[d][r] = ldg_stream(ap[r] + kk);
      // the two k-steps of a pair hit the same accumulator: issue every chain's first step, then every chain's second,
      // so that dependent DMMAs are NR * NB issues apart (dependent latency 26 cycles, issue 16)
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) dmma(acc[r][nb], a[r].x, b[nb][0]);
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) dmma(acc[r][nb], a[r].y, b[nb][1]);
    }
    bp += kRing * 64;
    if constexpr (PB) bp2 += kRing * 64;
  }
#pragma unroll
  for (int d = 0; d < kRing - 1; ++d)
    if (kp + d < cnt) {
      double b0[NB], b1[NB];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        b0[nb] = bp[nb * nbs + d * 64]; b1[nb] = bp[nb * nbs + d * 64 + 32];
        if constexpr (PB) {
          b0[nb] = 8 * (kp + d) < rowlim ? b0[nb] - bp2[nb * nbs + d * 64] : 0.0;
          b1[nb] = 8 * (kp + d) + 4 < rowlim ? b1[nb] - bp2[nb * nbs + d * 64 + 32] : 0.0;
        }
      }
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int r = 0; r < NR; ++r) dmma(acc[r][nb], ring[d][r].x, b0[nb]);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int r = 0; r < NR; ++r) dmma(acc[r][nb], ring[d][r].y, b1[nb]);
    }
}
// nr (1..kRG) row-blocks starting at rb: dispatch to the compile-time variants (nr is warp-uniform)
template <int NB, int RG, int kRing, bool PB = false>
__device__ __forceinline__ void gemm_seg(const double2 *__restrict__ opl, int kpt, int rb, int nr, int kp0, int cnt,
                                         const double *bp, int nbs, double (&acc)[RG][NB][2], const double *bp2 = nullptr, int rowlim = 0) {
  if (cnt <= 0) return;
  if constexpr (RG >= 4) {
    if (nr >= 4) { gemm_run<NB, 4, RG, kRing, PB>(opl, kpt, rb, kp0, cnt, bp, nbs, acc, bp2, rowlim); return; }
    if (nr == 3) { gemm_run<NB, 3, RG, kRing, PB>(opl, kpt, rb, kp0, cnt, bp, nbs, acc, bp2, rowlim); return; }
  }
  if (nr >= 2) gemm_run<NB, 2, RG, kRing, PB>(opl, kpt, rb, kp0, cnt, bp, nbs, acc, bp2, rowlim);
  else gemm_run<NB, 1, RG, kRing, PB>(opl, kpt, rb, kp0, cnt, bp, nbs, acc, bp2, rowlim);
}

// max over the 8 row-groups of a warp (lanes with equal lane&3 hold the same slot pair)
__device__ __forceinline__ double rmax8(double v) {
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 4));
  v = fmax(v, __shfl_xor_sync(0xffffffffu, v, 8));
  return fmax(v, __shfl_xor_sync(0xffffffffu, v, 16));
}
__device__ __forceinline__ double rsum8(double v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v + __shfl_xor_sync(0xffffffffu, v, 16);
}

}  // namespace

enum PassId { P_XBAR, P_PX, P_ATY, P_AX, P_ATD, P_DX, P_PD, P_ADX, P_QH, P_COUNT };

// WARPS = 8: two warps per SM sub-partition, up to 4 row-blocks per warp at once (254 registers);
// WARPS = 16: four warps per sub-partition, up to 2 row-blocks at once (128 registers): more warps to cover LDS / L2 / barrier stalls
// XD ("x-space, diagonal"): the rows come as [G; -G] with G = diag(a) after scaling (a two-sided box on the variables, e.g. the
// input limits of the condensed multi-input MPC, BASELINE config 3).  Then W = A̅V is dense for no reason: the iteration runs in
// x-space instead,   x~ = V (dinv .* (V' rhs)),   rhs = sigma x - q̄ + a .* wd,   z~_top = a .* x~,
// two n x n GEMMs (2 n^2 MACs per instance-iteration instead of n^2 + 2 m n = 3 n^2 with pairs), every product with A̅ or A̅' is
// element-wise in an epilogue, and a check needs ONE panel GEMM (P̄ x) instead of four.  Panels: cv's first n8 rows hold x̄, Sp the
// right-hand side between events, qh holds q̄.
// CTAS = 2 (with WARPS = 8): two independent tiles share an SM (128 registers per thread, 2 row-blocks per warp at once): their
// barriers and events are not synchronised, so one tile's check / store / refill overlaps the other's GEMMs on the DMMA pipe.
template <int NB, bool PAIRED, int WARPS, bool XD = false, int CTAS = 1>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
admm_shared_tile_kernel(TilePackDev K, SharedPlanDev P, BatchDev Bt, SettingsDev S, int *queue) {
  static_assert(!XD || PAIRED, "the x-space variant is for paired rows");
  constexpr int TB = 8 * NB;
  constexpr int kTileWarps = WARPS, kTileThreads = WARPS * 32, kRG = (WARPS == 8 && CTAS == 1) ? 4 : 2;   // (shadow the file-level defaults)
#ifdef SMPC_TILE_RING
  constexpr int kRing = SMPC_TILE_RING;
#else
  // operator prefetch distance in k-pairs (sweep in the file header): 2 for the x-space variant (config 3: 94.2 ms against 95.4 with 3),
  // 1 with two tiles per SM (config 5: the other tile covers the latency; 0.671 s per 100 steps against 0.683), 3 otherwise
  constexpr int kRing = XD ? 2 : (CTAS == 2 ? 1 : 3);
#endif
  extern __shared__ __align__(16) double smem[];
  const int n = P.n, m = P.m, n8 = K.n8, m8 = K.m8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q2 = 2 * (lane & 3);
  // panels
  // cv = [xi; w] as ONE panel of n8 + m8 rows per 8-slot block (the K dimension of GEMM 1)
  const int cvr = n8 + m8;
  double *cv = smem, *Tp = cv + cvr * TB, *qh = Tp + n8 * TB, *dinv = qh + n8 * TB, *Sp = dinv + n8 * TB, *Dp = Sp + n8 * TB;
  double *zp = Dp + n8 * TB, *yp = zp + m8 * TB, *lbp = yp + m8 * TB, *ubp = lbp + m8 * TB;
  const int bfrag = (lane & 3) * 8 + ((lane >> 2) ^ swz(lane & 3));   // this lane's element of a B fragment (k-row lane & 3, slot lane >> 2)
  TileCtl<TB, WARPS> &C = *reinterpret_cast<TileCtl<TB, WARPS> *>(ubp + m8 * TB);
  const int NRB = n8 >> 3, MRB = m8 >> 3;
  const int nrb0 = (warp * NRB) / kTileWarps, nrb1 = ((warp + 1) * NRB) / kTileWarps;
  const int mrb0 = (warp * MRB) / kTileWarps, mrb1 = ((warp + 1) * MRB) / kTileWarps;
  const int kpN = n8 >> 3, kpM = m8 >> 3;
  // PAIRED: rows r and r + mp of W are negatives of each other; the iteration GEMMs use the top half only and the first
  // mp rows of the w part of the cv panel carry wd = w_top - w_bot between events (events see the full panels)
  const int mp = PAIRED ? K.mp : 0, kpMp = PAIRED ? (K.mp8 >> 3) : 0, MRBp = kpMp;
  const int prb0 = (warp * MRBp) / kTileWarps, prb1 = ((warp + 1) * MRBp) / kTileWarps;
  const double alpha = S.alpha, oma = 1.0 - S.alpha;
  const int check_every = S.check_every, adapt_every = (S.adaptive_rho && S.rho_interval > 0) ? S.rho_interval : 0;

  auto zero_acc = [&](double (&acc)[kRG][NB][2]) {
#pragma unroll
    for (int r = 0; r < kRG; ++r)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) acc[r][nb][0] = acc[r][nb][1] = 0.0;
  };
  auto pidx = [&](int rows, int nb, int row) { return (nb * rows + row) * 8 + (q2 ^ swz(row)); };   // this thread's double2 (slots q2, q2 + 1) in a panel
  auto xi_at = [&](int e_n) { const int nb = e_n / (n8 * 8); return e_n + nb * m8 * 8; };            // n-panel index -> cv index (xi part)
  auto w_at = [&](int e_m) { const int nb = e_m / (m8 * 8); return e_m + (nb + 1) * n8 * 8; };        // m-panel index -> cv index (w part)

  // ---- init: empty tile
  for (int e = tid; e < (6 * n8 + 5 * m8) * TB; e += kTileThreads) smem[e] = 0.0;
  if (tid < TB) {
    C.inst[tid] = -1; C.it0[tid] = 0; C.rho_up[tid] = 0; C.status[tid] = SMPC_UNSOLVED; C.flags[tid] = 0;
    C.rho[tid] = 1.0; C.rinv[tid] = 1.0; C.rho_eq[tid] = 1.0; C.rinv_eq[tid] = 1.0;
  }
  if (tid == 0) { C.next_event = 0x7fffffff; C.active = 0; C.refill_mask = (1 << TB) - 1; C.work_flag = 0; }
  __syncthreads();
  for (int e = tid; e < m8 * TB; e += kTileThreads) { lbp[e] = -1.0; ubp[e] = 1.0; }
  for (int e = tid; e < n8 * TB; e += kTileThreads) dinv[e] = 1.0;
  __syncthreads();

  int k = 0;       // tile iteration counter
  bool event = true, initial = true;

#ifdef SMPC_TILE_PROFILE
  long long pf_pass[P_COUNT + 3] = {0}, pf_pre[P_COUNT] = {0}, pf_it = 0, pf_t0 = clock64(), pf_c = 0, pf_x[8] = {0}, pf_i[4] = {0}; int pf_nev = 0;
#define PFM(arr, i) { const long long tt = clock64(); arr[i] += tt - pf_c; pf_c = tt; }
#else
#define PFM(arr, i)
#endif
  for (;;) {
#ifdef SMPC_TILE_PROFILE
    if (event) { const long long tt = clock64(); pf_it += tt - pf_t0; pf_c = tt; }
#endif
    if (event) {
      // ================= event: termination check / rho adaptation / max_iter for the slots that are due, then
      // store + refill of the slots that finished.  The first event (k = 0) only fills the tile.
      const double c = P.c, cinv = P.cinv;
      const bool unscale = !S.scaled_termination;
      const double qnan = __longlong_as_double(0x7ff8000000000000LL);
      const bool warm = S.warm_start && !Bt.fresh;
      auto nrm = [&](int id, int s) { return __longlong_as_double((long long)C.nmax[id][s]); };
      auto total = [&](int id, int s) { double a = 0.0; for (int w = 0; w < kTileWarps; ++w) a += C.psum[id][w][s]; return a; };
      auto any_flags = [&]() { int a = 0; for (int s = 0; s < TB; ++s) a |= C.flags[s]; return a; };

      for (int e = tid; e < N_COUNT * TB; e += kTileThreads) (&C.nmax[0][0])[e] = 0ull;
      for (int e = tid; e < S_COUNT * kTileWarps * TB; e += kTileThreads) (&C.psum[0][0][0])[e] = 0.0;
      if (!initial) {
        // The slots that finish at this event are refilled ~10^5 cycles from now with the next QPs of the queue.  Their rows
        // (q, l, u and the warm-start state) are cold in HBM: pull the next 2 TB instances of the queue into L2 now (whichever
        // tile pops them finds them there).  Speculative and read-only: results do not depend on it.
        const int head = *reinterpret_cast<volatile int *>(queue);
        const int lines_n = (n * 8 + 127) >> 7, lines_m = (m * 8 + 127) >> 7;          // 128-byte lines per row vector
        const int per = (Bt.q ? lines_n : 0) + (Bt.l ? lines_m : 0) + (Bt.u ? lines_m : 0) + (warm ? lines_n + 2 * lines_m : 0);
        for (int e = tid; e < 2 * TB * per; e += kTileThreads) {
          const int b = head + e / per;
          if (b >= Bt.B) break;
          int j = e % per;
          const double *p = nullptr;
          if (Bt.q) { if (j < lines_n) p = Bt.q + (size_t)b * n + 16 * j; j -= lines_n; }
          if (!p && Bt.l) { if (j < lines_m) p = Bt.l + (size_t)b * m + 16 * j; j -= lines_m; }
          if (!p && Bt.u) { if (j < lines_m) p = Bt.u + (size_t)b * m + 16 * j; j -= lines_m; }
          if (!p && warm) {
            if (j < lines_n) p = Bt.xi + (size_t)b * n + 16 * j;
            else if (j < lines_n + lines_m) p = Bt.z + (size_t)b * m + 16 * (j - lines_n);
            else p = Bt.y + (size_t)b * m + 16 * (j - lines_n - lines_m);
          }
          if (p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
      }
      if (tid < TB) {
        int f = 0;
        if (C.inst[tid] >= 0) {
          const int li = k - C.it0[tid];
          if (check_every > 0 && li % check_every == 0) f |= F_DUE_CHECK | F_CADENCE;
          if (adapt_every > 0 && li % adapt_every == 0) f |= F_DUE_ADAPT;
          if (li >= S.max_iter) f |= F_AT_MAX | F_DUE_CHECK;
        }
        C.flags[tid] = f;
      }
      __syncthreads();
      int any = initial ? 0 : any_flags();
      int done = 0;

      PFM(pf_pass, P_COUNT)
      for (int pass = 0; pass < P_COUNT; ++pass) {
        // ---------- what runs before the GEMM of this pass (block-uniform control flow)
        bool run = false;
        if (pass <= P_AX) {
          run = !initial;
          if (XD && pass == P_XBAR) {   // x̄ is the iterate itself: copy it where the later passes and the store expect it
            run = false;
            if (!initial) {
              for (int e = tid; e < n8 * TB; e += kTileThreads) Tp[e] = cv[xi_at(e)];
              __syncthreads();
            }
          }
        } else if (pass == P_ATD) {
          if (!initial) {
            __syncthreads();   // norms of the four update_info passes are complete
            // per-slot: update_info results, convergence tests with the plain tolerances
            if (tid < TB && C.flags[tid]) {
              const int s = tid;
              int f = C.flags[s];
              const double pri = m == 0 ? 0.0 : (unscale ? nrm(N_RP_U, s) : nrm(N_RP_S, s));
              const double dua = unscale ? cinv * nrm(N_RD_U, s) : nrm(N_RD_S, s);
              C.pri[s] = pri; C.dua[s] = dua;
              C.obj[s] = (unscale ? cinv : 1.0) * total(S_OBJ, s);
              if (f & F_DUE_CHECK) {
                const double nEz = unscale ? nrm(N_Z_U, s) : nrm(N_Z_S, s), nEAx = unscale ? nrm(N_AX_U, s) : nrm(N_AX_S, s);
                const double nDq = unscale ? nrm(N_Q_U, s) : nrm(N_Q_S, s), nDAty = unscale ? nrm(N_ATY_U, s) : nrm(N_ATY_S, s);
                const double nDPx = unscale ? nrm(N_PX_U, s) : nrm(N_PX_S, s);
                if (m == 0 || pri < S.eps_abs + S.eps_rel * fmax(nEz, nEAx)) f |= F_PRIM_OK; else f |= F_NEED_PINF;
                if (dua < S.eps_abs + S.eps_rel * (unscale ? cinv : 1.0) * fmax(fmax(nDq, nDAty), nDPx)) f |= F_DUAL_OK; else f |= F_NEED_DINF;
              }
              C.flags[s] = f;
            }
            __syncthreads();
            any = any_flags();
            // is_primal_infeasible: project delta_y (carried by the w panel) on the polar of the recession cone
            if (any & F_NEED_PINF) {
              double lh[NB][2];
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) lh[nb][0] = lh[nb][1] = 0.0;
              for (int rb = mrb0; rb < mrb1; ++rb) {
                const int row = 8 * rb + g;
                const double E = row < m ? __ldg(P.E + row) : 1.0;
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                  const int pi = pidx(m8, nb, row);
                  const double2 d = *reinterpret_cast<const double2 *>(cv + pidx(cvr, nb, n8 + row));
                  const double2 lo = *reinterpret_cast<const double2 *>(lbp + pi), hi = *reinterpret_cast<const double2 *>(ubp + pi);
                  double dd[2] = {d.x, d.y}, nd[2];
                  const double lo2[2] = {lo.x, lo.y}, hi2[2] = {hi.x, hi.y};
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    const bool uinf = hi2[j] > kInfty * kMinScaling, linf = lo2[j] < -kInfty * kMinScaling;
                    if (uinf) dd[j] = linf ? 0.0 : fmin(dd[j], 0.0); else if (linf) dd[j] = fmax(dd[j], 0.0);
                    if (row >= m) dd[j] = 0.0;
                    nd[j] = fabs(unscale ? E * dd[j] : dd[j]);
                    const double dp = fmax(dd[j], 0.0), dm = fmin(dd[j], 0.0);
                    if (dp != 0.0) lh[nb][j] += hi2[j] * dp;
                    if (dm != 0.0) lh[nb][j] += lo2[j] * dm;
                  }
                  *reinterpret_cast<double2 *>(cv + pidx(cvr, nb, n8 + row)) = make_double2(dd[0], dd[1]);
                  const double v0 = rmax8(nd[0]), v1 = rmax8(nd[1]);
                  if (g == 0) {
                    atomicMax(&C.nmax[N_DY][nb * 8 + q2], (unsigned long long)__double_as_longlong(v0));
                    atomicMax(&C.nmax[N_DY][nb * 8 + q2 + 1], (unsigned long long)__double_as_longlong(v1));
                  }
                }
              }
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) {
                const double s0 = rsum8(lh[nb][0]), s1 = rsum8(lh[nb][1]);
                if (g == 0) { C.psum[S_LHS][warp][nb * 8 + q2] = s0; C.psum[S_LHS][warp][nb * 8 + q2 + 1] = s1; }
              }
              __syncthreads();
              if (tid < TB && (C.flags[tid] & F_NEED_PINF)) {
                const double nd = nrm(N_DY, tid), lhs = total(S_LHS, tid);
                if (nd > S.eps_prim_inf && lhs < -S.eps_prim_inf * nd) C.flags[tid] |= F_NEED_ATD;
              }
              __syncthreads();
              any = any_flags();
            }
            run = (any & F_NEED_ATD) != 0;
          }
        } else if (pass == P_DX) {
          // delta_xi of the event iteration is in the Dp panel.  Exact screen before the delta_x = V delta_xi GEMM:
          // ||D delta_x||_inf <= ||D V||_inf ||delta_xi||_inf, so a slot whose bound is below eps_dual_inf cannot pass
          // is_dual_infeasible's first test (at either tolerance) and needs none of the three dual-infeasibility passes
          if (any & F_NEED_DINF) {
            for (int rb = nrb0; rb < nrb1; ++rb) {
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) {
                const double2 d = *reinterpret_cast<const double2 *>(Dp + pidx(n8, nb, 8 * rb + g));
                const double v0 = rmax8(fabs(d.x)), v1 = rmax8(fabs(d.y));
                if (g == 0) {
                  atomicMax(&C.nmax[N_DXI][nb * 8 + q2], (unsigned long long)__double_as_longlong(v0));
                  atomicMax(&C.nmax[N_DXI][nb * 8 + q2 + 1], (unsigned long long)__double_as_longlong(v1));
                }
              }
            }
            __syncthreads();
            if (tid < TB && (C.flags[tid] & F_NEED_DINF) && 1.000001 * (XD ? K.dmax : P.dx_bound) * nrm(N_DXI, tid) <= S.eps_dual_inf) C.flags[tid] &= ~F_NEED_DINF;
            __syncthreads();
            any = any_flags();
          }
          run = (any & F_NEED_DINF) != 0;
        } else if (pass == P_PD) {
          if (any & F_NEED_DINF) {
            __syncthreads();
            if (tid < TB && (C.flags[tid] & F_NEED_DINF)) {
              const double nd = nrm(N_DX, tid), qd = total(S_QD, tid), cs = unscale ? c : 1.0;
              if (nd > S.eps_dual_inf && qd < -cs * S.eps_dual_inf * nd) C.flags[tid] |= F_NEED_DINF2;
            }
            __syncthreads();
            any = any_flags();
          }
          run = (any & F_NEED_DINF2) != 0;
        } else if (pass == P_ADX) {
          // OSQP's third condition, ||P dx|| < eps ||dx||, is known now: a slot that fails it even at the 10 x tolerance of the
          // approximate check cannot be dual infeasible, and A̅ dx (this pass) is needed by nobody else
          if (any & F_NEED_DINF2) {
            __syncthreads();   // the norms of the P̄ dx pass are complete
            if (tid < TB && (C.flags[tid] & F_NEED_DINF2)) {
              const double nd = nrm(N_DX, tid), cs = unscale ? c : 1.0;
              if (!(nrm(N_PD, tid) < cs * 10.0 * S.eps_dual_inf * nd)) C.flags[tid] &= ~F_NEED_DINF2;
            }
            __syncthreads();
            any = any_flags();
          }
          run = (any & F_NEED_DINF2) != 0;
        } else {   // P_QH: decisions, store, refill, then q̂ for the (re)filled tile
          __syncthreads();
          if (!initial && tid < TB && C.flags[tid]) {
            const int s = tid;
            int f = C.flags[s];
            int status = SMPC_UNSOLVED;
            double obj = C.obj[s];
            auto decide = [&](bool approx) {
              const double mul = approx ? 10.0 : 1.0;
              const double ea = mul * S.eps_abs, er = mul * S.eps_rel, epi = mul * S.eps_prim_inf, edi = mul * S.eps_dual_inf;
              const double nEz = unscale ? nrm(N_Z_U, s) : nrm(N_Z_S, s), nEAx = unscale ? nrm(N_AX_U, s) : nrm(N_AX_S, s);
              const double nDq = unscale ? nrm(N_Q_U, s) : nrm(N_Q_S, s), nDAty = unscale ? nrm(N_ATY_U, s) : nrm(N_ATY_S, s);
              const double nDPx = unscale ? nrm(N_PX_U, s) : nrm(N_PX_S, s);
              bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
              if (m == 0 || C.pri[s] < ea + er * fmax(nEz, nEAx)) prim_ok = true;
              else {
                const double nd = nrm(N_DY, s), lhs = total(S_LHS, s);
                if (nd > epi && lhs < -epi * nd) prim_inf = nrm(N_ATD, s) < epi * nd;
              }
              if (C.dua[s] < ea + er * (unscale ? cinv : 1.0) * fmax(fmax(nDq, nDAty), nDPx)) dual_ok = true;
              else {
                const double nd = nrm(N_DX, s), qd = total(S_QD, s), cs = unscale ? c : 1.0;
                if (nd > edi && qd < -cs * edi * nd && nrm(N_PD, s) < cs * edi * nd) dual_inf = !(f & (approx ? (F_BADROW << 16) : F_BADROW));
              }
              if (prim_ok && dual_ok) status = approx ? SMPC_SOLVED_INACCURATE : SMPC_SOLVED;
              else if (prim_inf) { status = approx ? SMPC_PRIMAL_INFEASIBLE_INACCURATE : SMPC_PRIMAL_INFEASIBLE; obj = kInfty; }
              else if (dual_inf) { status = approx ? SMPC_DUAL_INFEASIBLE_INACCURATE : SMPC_DUAL_INFEASIBLE; obj = -kInfty; }
            };
            // order of admm_shared_generic.cu: cadence check, rho adaptation, then (at max_iter) the final plain + 10x checks
            if (f & F_CADENCE) decide(false);
            if ((f & F_DUE_ADAPT) && status == SMPC_UNSOLVED) {
              // compute_rho_estimate / adapt_rho on the SCALED norms
              const double rho = C.rho[s];
              const double pr = nrm(N_RP_S, s) / (fmax(nrm(N_Z_S, s), nrm(N_AX_S, s)) + kDivTol);
              const double dr = nrm(N_RD_S, s) / (fmax(fmax(nrm(N_Q_S, s), nrm(N_ATY_S, s)), nrm(N_PX_S, s)) + kDivTol);
              const double rn = fmin(fmax(rho * sqrt(pr / (dr + kDivTol)), kRhoMin), kRhoMax);
              if (rn > rho * S.rho_tol || rn < rho / S.rho_tol) {
                C.rho[s] = rn; C.rinv[s] = 1.0 / rn; C.rho_eq[s] = kRhoEqOverIneq * rn; C.rinv_eq[s] = 1.0 / (kRhoEqOverIneq * rn);
                C.rho_up[s]++; f |= F_RHO_NEW;
              }
            }
            if ((f & F_AT_MAX) && status == SMPC_UNSOLVED) {
              if (!(f & F_CADENCE)) decide(false);
              if (status == SMPC_UNSOLVED) {
                decide(true);
                if (status == SMPC_UNSOLVED) status = SMPC_MAX_ITER_REACHED;
              }
            }
            if (status != SMPC_UNSOLVED) { f |= F_DONE; C.status[s] = status; C.obj[s] = obj; }
            C.flags[s] = f;
          }
          __syncthreads();
          PFM(pf_x, 0)
          // ---- store_solution for the QPs that finished; new dinv where rho changed
          int rho_new = 0;
          if (!initial)
            for (int s = 0; s < TB; ++s) { if (C.flags[s] & F_DONE) done |= 1 << s; if (C.flags[s] & F_RHO_NEW) rho_new |= 1 << s; }
          if (done | rho_new) {
            for (int nb = 0, s8 = tid & 7; nb < NB; ++nb)
_Pragma("unroll 4")
            for (int i = tid >> 3; i < n8; i += kTileThreads / 8) {
              const int e = (nb * n8 + i) * 8 + s8, s = nb * 8 + (s8 ^ swz(i));   // s8 = position in the row
              if (((rho_new & ~done) >> s) & 1) dinv[e] = 1.0 / (1.0 + C.rho[s] * (i < n ? __ldg(P.lam + i) : 0.0));
              if (!((done >> s) & 1)) continue;
              if (i < n) {
                const int b = C.inst[s], st = C.status[s];
                const bool has_sol = !(st == SMPC_PRIMAL_INFEASIBLE || st == SMPC_PRIMAL_INFEASIBLE_INACCURATE || st == SMPC_DUAL_INFEASIBLE || st == SMPC_DUAL_INFEASIBLE_INACCURATE);
                if (Bt.x_out) Bt.x_out[(size_t)b * n + i] = has_sol ? __ldg(P.D + i) * Tp[e] : qnan;
                Bt.xi[(size_t)b * n + i] = has_sol ? cv[xi_at(e)] : 0.0;
              }
              cv[xi_at(e)] = 0.0; Dp[e] = 0.0; Tp[e] = 0.0; qh[e] = 0.0; dinv[e] = 1.0;   // an unfilled slot iterates on zeros
            }
          }
          if (done) {
            for (int nb = 0, s8 = tid & 7; nb < NB; ++nb)
#pragma unroll kUnrollSt
            for (int r = tid >> 3; r < m8; r += kTileThreads / 8) {
              const int e = (nb * m8 + r) * 8 + s8, s = nb * 8 + (s8 ^ swz(r));
              if (!((done >> s) & 1)) continue;
              if (r < m) {
                const int b = C.inst[s], st = C.status[s];
                const bool has_sol = !(st == SMPC_PRIMAL_INFEASIBLE || st == SMPC_PRIMAL_INFEASIBLE_INACCURATE || st == SMPC_DUAL_INFEASIBLE || st == SMPC_DUAL_INFEASIBLE_INACCURATE);
                if (Bt.y_out) Bt.y_out[(size_t)b * m + r] = has_sol ? cinv * (__ldg(P.E + r) * yp[e]) : qnan;
                Bt.z[(size_t)b * m + r] = has_sol ? zp[e] : 0.0;
                Bt.y[(size_t)b * m + r] = has_sol ? yp[e] : 0.0;
              }
              zp[e] = 0.0; yp[e] = 0.0; lbp[e] = -1.0; ubp[e] = 1.0;
            }
            if (tid < TB && ((done >> tid) & 1)) {
              const int b = C.inst[tid];
              Bt.rho[b] = C.rho[tid]; Bt.status[b] = C.status[tid]; Bt.iter[b] = k - C.it0[tid]; Bt.rho_updates[b] = C.rho_up[tid];
              Bt.obj[b] = C.obj[tid]; Bt.pri_res[b] = C.pri[tid]; Bt.dua_res[b] = C.dua[tid];
            }
            if (tid == 0) C.refill_mask = done;
          }
          __syncthreads();
          PFM(pf_x, 1)
          // ---- refill: pull new QPs from the queue into the slots of C.refill_mask.  QPs with invalid bounds (see
          // admm_shared_generic.cu) are stored as UNSOLVED at once and their slot is refilled again.
          const bool any_refill = C.refill_mask != 0;
          while (any_refill) {
            // one ticket request per tile for all its free slots (the tiles reach their events in waves: per-slot requests
            // queue several hundred same-address atomics at the L2)
            int rm = 0, base = 0;
            if (tid < 32) {
              rm = C.refill_mask;
              if (tid == 0) base = atomicAdd(queue, __popc(rm));
              base = __shfl_sync(0xffffffffu, base, 0);
            }
            if (tid < TB && ((rm >> tid) & 1)) {
              const int b = base + __popc(rm & ((1 << tid) - 1));
              SMPC_DBG(b >= 0, "tile queue ticket");
              const bool ok = b < Bt.B;
              C.inst[tid] = ok ? b : -1;
              C.it0[tid] = k; C.rho_up[tid] = 0; C.status[tid] = SMPC_UNSOLVED; C.flags[tid] = 0;
              const double rho = ok ? (Bt.fresh ? fmin(fmax(S.rho0, kRhoMin), kRhoMax) : Bt.rho[b]) : 1.0;
              C.rho[tid] = rho; C.rinv[tid] = 1.0 / rho; C.rho_eq[tid] = kRhoEqOverIneq * rho; C.rinv_eq[tid] = 1.0 / (kRhoEqOverIneq * rho);
              C.obj[tid] = 0.0; C.pri[tid] = 0.0; C.dua[tid] = 0.0;
            }
            __syncthreads();
            PFM(pf_x, 2)
            const int mask = C.refill_mask;
            // slot-fast mapping (lane & 7 = slot): conflict-free panel accesses
            for (int nb = 0, s8 = tid & 7; nb < NB; ++nb)
_Pragma("unroll 4")
            for (int i = tid >> 3; i < n8; i += kTileThreads / 8) {
              const int e = (nb * n8 + i) * 8 + s8, s = nb * 8 + (s8 ^ swz(i));   // s8 = position in the row
              if (!((mask >> s) & 1)) continue;
              const int b = C.inst[s];
              const double v = (b >= 0 && i < n && warm) ? Bt.xi[(size_t)b * n + i] : 0.0;
              cv[xi_at(e)] = v; Dp[e] = 0.0;
              dinv[e] = b >= 0 ? 1.0 / (1.0 + C.rho[s] * (i < n ? __ldg(P.lam + i) : 0.0)) : 1.0;
            }
            for (int nb = 0, s8 = tid & 7; nb < NB; ++nb)
#pragma unroll kUnrollLd
            for (int r = tid >> 3; r < m8; r += kTileThreads / 8) {
              const int e = (nb * m8 + r) * 8 + s8, s = nb * 8 + (s8 ^ swz(r));
              if (!((mask >> s) & 1)) continue;
              const int b = C.inst[s];
              double lo = -1.0, hi = 1.0, zz = 0.0, yy = 0.0;
              if (b >= 0 && r < m) {
                const double E = __ldg(P.E + r);
                lo = E * (Bt.l ? Bt.l[(size_t)b * m + r] : __ldg(P.l0 + r));
                hi = E * (Bt.u ? Bt.u[(size_t)b * m + r] : __ldg(P.u0 + r));
                if (warm) { zz = Bt.z[(size_t)b * m + r]; yy = Bt.y[(size_t)b * m + r]; }
                if (lo > hi) atomicOr(&C.flags[s], F_BADBOUNDS);   // (a class change keeps the plan's rho_vec entry: admm_shared_generic.cu)
              }
              lbp[e] = lo; ubp[e] = hi; zp[e] = zz; yp[e] = yy;
            }
            __syncthreads();
            PFM(pf_x, 3)
            int again = 0;
            for (int s = 0; s < TB; ++s) {
              if (!((mask >> s) & 1) || C.inst[s] < 0 || !(C.flags[s] & F_BADBOUNDS)) continue;
              again |= 1 << s;
              const int b = C.inst[s];
              for (int i = tid; i < n; i += kTileThreads) { if (Bt.x_out) Bt.x_out[(size_t)b * n + i] = qnan; Bt.xi[(size_t)b * n + i] = 0.0; }
              for (int r = tid; r < m; r += kTileThreads) {
                if (Bt.y_out) Bt.y_out[(size_t)b * m + r] = qnan;
                Bt.z[(size_t)b * m + r] = 0.0; Bt.y[(size_t)b * m + r] = 0.0;
              }
              if (tid == 0) {
                Bt.rho[b] = C.rho[s]; Bt.status[b] = SMPC_UNSOLVED; Bt.iter[b] = 0; Bt.rho_updates[b] = 0;
                Bt.obj[b] = 0.0; Bt.pri_res[b] = 0.0; Bt.dua_res[b] = 0.0;
              }
            }
            __syncthreads();
            PFM(pf_x, 4)
            if (!again) break;
            if (tid == 0) C.refill_mask = again;
            for (int nb = 0, s8 = tid & 7; nb < NB; ++nb)
_Pragma("unroll 4")
            for (int r = tid >> 3; r < m8; r += kTileThreads / 8) {
              const int e = (nb * m8 + r) * 8 + s8, s = nb * 8 + (s8 ^ swz(r));
              if ((again >> s) & 1) { lbp[e] = -1.0; ubp[e] = 1.0; zp[e] = 0.0; yp[e] = 0.0; }
            }
            __syncthreads();
          }
          if (any_refill) {
            // q̄ = c D q (osqp_update_lin_cost) of every slot -> S ; q̂ = V' q̄ follows as this pass's GEMM
            for (int nb = 0, s8 = tid & 7; nb < NB; ++nb)
_Pragma("unroll 4")
            for (int i = tid >> 3; i < n8; i += kTileThreads / 8) {
              const int e = (nb * n8 + i) * 8 + s8, s = nb * 8 + (s8 ^ swz(i));   // s8 = position in the row
              const int b = C.inst[s];
              Sp[e] = (b >= 0 && i < n && Bt.q) ? c * (__ldg(P.D + i) * Bt.q[(size_t)b * n + i]) : 0.0;
            }
            __syncthreads();
            run = true;
          }
        }
        PFM(pf_pre, pass)
        if (!run) continue;

        // ---------- the GEMM of this pass: one loop for every operator / panel combination
        const double2 *op; const double *panel; int kpt, rb0, rb1, krows;   // krows: panel rows per 8-slot block
        switch (pass) {
          case P_XBAR: op = reinterpret_cast<const double2 *>(K.Vp);  kpt = kpN; panel = cv;          krows = cvr; rb0 = nrb0; rb1 = nrb1; break;
          case P_PX:   op = reinterpret_cast<const double2 *>(XD ? K.Pp : K.PVp); kpt = kpN; panel = cv; krows = cvr; rb0 = nrb0; rb1 = nrb1; break;
          case P_ATY:  op = reinterpret_cast<const double2 *>(K.ATp); kpt = kpM; panel = yp;          krows = m8;  rb0 = nrb0; rb1 = nrb1; break;
          case P_AX:   op = reinterpret_cast<const double2 *>(K.Wp);  kpt = kpN; panel = cv;          krows = cvr; rb0 = mrb0; rb1 = mrb1;
                       // row pairs: A̅ x̄ of the top half only (row p + mp is its negative): half the DMMAs of this pass
                       if (PAIRED && !XD) { op = reinterpret_cast<const double2 *>(K.Wtop); rb0 = prb0; rb1 = prb1; }
                       break;
          case P_ATD:  op = reinterpret_cast<const double2 *>(K.ATp); kpt = kpM; panel = cv + n8 * 8; krows = cvr; rb0 = nrb0; rb1 = nrb1; break;
          case P_DX:   op = reinterpret_cast<const double2 *>(K.Vp);  kpt = kpN; panel = Dp;          krows = n8;  rb0 = nrb0; rb1 = nrb1; break;
          case P_PD:   op = reinterpret_cast<const double2 *>(XD ? K.Pp : K.PVp); kpt = kpN; panel = Dp; krows = n8;  rb0 = nrb0; rb1 = nrb1; break;
          case P_ADX:  op = reinterpret_cast<const double2 *>(K.Wp);  kpt = kpN; panel = Dp;          krows = n8;  rb0 = mrb0; rb1 = mrb1; break;
          default:     op = reinterpret_cast<const double2 *>(K.VTp); kpt = kpN; panel = Sp;          krows = n8;  rb0 = nrb0; rb1 = nrb1; break;
        }
        const int nmax_base = pass == P_ATY ? N_RD_S : pass == P_AX ? N_RP_S : pass == P_ATD ? N_ATD : pass == P_DX ? N_DX : N_PD;
        const int nmax_cnt = pass == P_ATY ? 8 : pass == P_AX ? 6 : (pass == P_ATD || pass == P_DX || pass == P_PD) ? 1 : 0;
        const int sum_id = pass == P_ATY ? S_OBJ : pass == P_DX ? S_QD : -1;
        for (int rb = rb0; rb < rb1; rb += kRG) {
          double acc[kRG][NB][2];
          zero_acc(acc);
          if (XD && pass != P_PX && pass != P_PD) {
            // products with A̅ = [diag(a); -diag(a)] (or the identity) are element-wise: fill the accumulators directly
#pragma unroll
            for (int r = 0; r < kRG; ++r) {
              if (rb + r >= rb1) continue;
              const int row = 8 * (rb + r) + g;
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) {
                double2 v = make_double2(0.0, 0.0);
                if (pass == P_ATY || pass == P_ATD) {            // A̅' y = a .* (y_top - y_bot); A̅' delta_y likewise (delta_y is in the w part of cv)
                  if (row < n) {
                    const double a = __ldg(K.adiag + row);
                    const double2 t = pass == P_ATY ? *reinterpret_cast<const double2 *>(yp + pidx(m8, nb, row)) : *reinterpret_cast<const double2 *>(cv + pidx(cvr, nb, n8 + row));
                    const double2 b = pass == P_ATY ? *reinterpret_cast<const double2 *>(yp + pidx(m8, nb, row + mp)) : *reinterpret_cast<const double2 *>(cv + pidx(cvr, nb, n8 + row + mp));
                    v = make_double2(a * (t.x - b.x), a * (t.y - b.y));
                  }
                } else if (pass == P_AX || pass == P_ADX) {      // A̅ x̄ (rows over m): +a x on the top half, -a x on the bottom half
                  if (row < m) {
                    const int col = row < mp ? row : row - mp;
                    const double a = row < mp ? __ldg(K.adiag + col) : -__ldg(K.adiag + col);
                    const double2 x2 = pass == P_AX ? *reinterpret_cast<const double2 *>(cv + pidx(cvr, nb, col)) : *reinterpret_cast<const double2 *>(Dp + pidx(n8, nb, col));
                    v = make_double2(a * x2.x, a * x2.y);
                  }
                } else if (pass == P_DX) {                       // delta_x is the Dp panel itself
                  v = *reinterpret_cast<const double2 *>(Dp + pidx(n8, nb, row));
                } else {                                         // P_QH: q̄ itself
                  v = *reinterpret_cast<const double2 *>(Sp + pidx(n8, nb, row));
                }
                acc[r][nb][0] = v.x; acc[r][nb][1] = v.y;
              }
            }
          } else if (PAIRED && !XD && pass == P_ATY) {
            // row pairs: A̅'y = G'(y_top - y_bot): half the DMMAs and operator bytes of this pass (K = mp instead of m); the B fragment is the
            // difference of panel rows r and r + mp (the swizzle of a thread's rows does not change from k-step to k-step: 4 rows apart)
            const int bfrag2 = (mp + (lane & 3)) * 8 + ((lane >> 2) ^ swz(mp + (lane & 3)));
            gemm_seg<NB, kRG, kRing, true>(reinterpret_cast<const double2 *>(K.ATtop) + lane, kpMp, rb, min(kRG, rb1 - rb), 0, kpMp, yp + bfrag, m8 * 8, acc, yp + bfrag2, mp - (lane & 3));
          } else {
            gemm_seg<NB, kRG, kRing>(op + lane, kpt, rb, min(kRG, rb1 - rb), 0, kpt, panel + bfrag, krows * 8, acc);
          }
          double mx[NB][8][2], sm[NB][2];
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            sm[nb][0] = sm[nb][1] = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) mx[nb][j][0] = mx[nb][j][1] = 0.0;
          }
#pragma unroll
          for (int r = 0; r < kRG; ++r) {
            if (rb + r >= rb1) continue;
            const int row = 8 * (rb + r) + g;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
              const double a0 = acc[r][nb][0], a1 = acc[r][nb][1];
              if (pass == P_XBAR) *reinterpret_cast<double2 *>(Tp + pidx(n8, nb, row)) = make_double2(a0, a1);
              else if (pass == P_PX) *reinterpret_cast<double2 *>(Sp + pidx(n8, nb, row)) = make_double2(a0, a1);
              else if (pass == P_QH) *reinterpret_cast<double2 *>(qh + pidx(n8, nb, row)) = make_double2(a0, a1);
              else if (pass == P_ATY) {
                // dual residual P̄x + q̄ + A̅'y (the owner thread reads back its own x̄ / P̄x entries)
                const double Di = row < n ? __ldg(P.D + row) : 1.0, Dinv = row < n ? __ldg(P.Dinv + row) : 1.0;
                const int pi = pidx(n8, nb, row);
                const double2 xb = *reinterpret_cast<const double2 *>(Tp + pi), px = *reinterpret_cast<const double2 *>(Sp + pi);
                double2 qh2 = make_double2(0.0, 0.0);
                if (XD) qh2 = *reinterpret_cast<const double2 *>(qh + pi);   // x-space: the qh panel holds q̄ = c D q itself (0 for an empty slot)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const int b = XD ? -1 : C.inst[nb * 8 + q2 + j];
                  const double qb = XD ? (j ? qh2.y : qh2.x) : ((b >= 0 && row < n && Bt.q) ? c * (Di * Bt.q[(size_t)b * n + row]) : 0.0);
                  const double aty = j ? a1 : a0, pxj = j ? px.y : px.x, xbj = j ? xb.y : xb.x;
                  const double rd = (qb + pxj) + aty;
                  mx[nb][0][j] = fmax(mx[nb][0][j], fabs(rd)); mx[nb][1][j] = fmax(mx[nb][1][j], fabs(qb));
                  mx[nb][2][j] = fmax(mx[nb][2][j], fabs(aty)); mx[nb][3][j] = fmax(mx[nb][3][j], fabs(pxj));
                  mx[nb][4][j] = fmax(mx[nb][4][j], fabs(Dinv * rd)); mx[nb][5][j] = fmax(mx[nb][5][j], fabs(Dinv * qb));
                  mx[nb][6][j] = fmax(mx[nb][6][j], fabs(Dinv * aty)); mx[nb][7][j] = fmax(mx[nb][7][j], fabs(Dinv * pxj));
                  sm[nb][j] += 0.5 * xbj * pxj + qb * xbj;
                }
              } else if (pass == P_AX) {
                if (PAIRED && !XD) {   // the accumulator is (A̅ x̄) of top row `row`; the bottom row row + mp has the opposite sign
                  if (row < mp) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                      const int rr = row + half * mp;
                      const double Einv = __ldg(P.Einv + rr), sg = half ? -1.0 : 1.0;
                      const double2 zz = *reinterpret_cast<const double2 *>(zp + pidx(m8, nb, rr));
#pragma unroll
                      for (int j = 0; j < 2; ++j) {
                        const double Ax = sg * (j ? a1 : a0), zj = j ? zz.y : zz.x, rp = Ax - zj;
                        mx[nb][0][j] = fmax(mx[nb][0][j], fabs(rp)); mx[nb][1][j] = fmax(mx[nb][1][j], fabs(zj)); mx[nb][2][j] = fmax(mx[nb][2][j], fabs(Ax));
                        mx[nb][3][j] = fmax(mx[nb][3][j], fabs(Einv * rp)); mx[nb][4][j] = fmax(mx[nb][4][j], fabs(Einv * zj));
                        mx[nb][5][j] = fmax(mx[nb][5][j], fabs(Einv * Ax));
                      }
                    }
                  }
                } else {
                const double Einv = row < m ? __ldg(P.Einv + row) : 1.0;
                const double2 zz = *reinterpret_cast<const double2 *>(zp + pidx(m8, nb, row));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const double Ax = j ? a1 : a0, zj = j ? zz.y : zz.x, rp = Ax - zj;
                  mx[nb][0][j] = fmax(mx[nb][0][j], fabs(rp)); mx[nb][1][j] = fmax(mx[nb][1][j], fabs(zj)); mx[nb][2][j] = fmax(mx[nb][2][j], fabs(Ax));
                  mx[nb][3][j] = fmax(mx[nb][3][j], fabs(Einv * rp)); mx[nb][4][j] = fmax(mx[nb][4][j], fabs(Einv * zj));
                  mx[nb][5][j] = fmax(mx[nb][5][j], fabs(Einv * Ax));
                }
                }
              } else if (pass == P_ATD || pass == P_PD) {
                const double Dinv = (unscale && row < n) ? __ldg(P.Dinv + row) : 1.0;
                mx[nb][0][0] = fmax(mx[nb][0][0], fabs(Dinv * a0)); mx[nb][0][1] = fmax(mx[nb][0][1], fabs(Dinv * a1));
              } else if (pass == P_DX) {
                const double Di = row < n ? __ldg(P.D + row) : 1.0;
                double2 qh2 = make_double2(0.0, 0.0);
                if (XD) qh2 = *reinterpret_cast<const double2 *>(qh + pidx(n8, nb, row));
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const int b = XD ? -1 : C.inst[nb * 8 + q2 + j];
                  const double qb = XD ? (j ? qh2.y : qh2.x) : ((b >= 0 && row < n && Bt.q) ? c * (Di * Bt.q[(size_t)b * n + row]) : 0.0);
                  const double dx = j ? a1 : a0;
                  mx[nb][0][j] = fmax(mx[nb][0][j], fabs(unscale ? Di * dx : dx));
                  sm[nb][j] += qb * dx;
                }
              } else if (row < m) {   // P_ADX: A̅ dx against the finite bounds, plain and 10x tolerance
                const double Einv = unscale ? __ldg(P.Einv + row) : 1.0;
                const int pi = pidx(m8, nb, row);
                const double2 lo = *reinterpret_cast<const double2 *>(lbp + pi), hi = *reinterpret_cast<const double2 *>(ubp + pi);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const int s = nb * 8 + q2 + j;
                  const double v = Einv * (j ? a1 : a0), nd = nrm(N_DX, s), lj = j ? lo.y : lo.x, hj = j ? hi.y : hi.x;
                  const double e1 = S.eps_dual_inf * nd, e10 = 10.0 * e1;
                  const bool fu = hj < kInfty * kMinScaling, fl = lj > -kInfty * kMinScaling;
                  int bits = 0;
                  if ((fu && v > e1) || (fl && v < -e1)) bits |= F_BADROW;
                  if ((fu && v > e10) || (fl && v < -e10)) bits |= F_BADROW << 16;
                  if (bits) atomicOr(&C.flags[s], bits);
                }
              }
            }
          }
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < nmax_cnt) {
                const double v0 = rmax8(mx[nb][j][0]), v1 = rmax8(mx[nb][j][1]);
                if (g == 0) {
                  atomicMax(&C.nmax[nmax_base + j][nb * 8 + q2], (unsigned long long)__double_as_longlong(v0));
                  atomicMax(&C.nmax[nmax_base + j][nb * 8 + q2 + 1], (unsigned long long)__double_as_longlong(v1));
                }
              }
            if (sum_id >= 0) {
              const double s0 = rsum8(sm[nb][0]), s1 = rsum8(sm[nb][1]);
              if (g == 0) { C.psum[sum_id][warp][nb * 8 + q2] += s0; C.psum[sum_id][warp][nb * 8 + q2 + 1] += s1; }
            }
          }
        }
        PFM(pf_pass, pass)
      }

      PFM(pf_pass, P_COUNT + 1)
      // ---- w = rho_vec z - y (the w panel carried delta_y); schedule the next event
      __syncthreads();
      {
        const int s8 = tid & 7;                      // thread -> position inside a row of an 8-slot block, rows tid / 8 + 32 j (no divisions)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          if (XD) {   // rhs = sigma x - q̄ + a .* ((rho_vec z - y)_top - (rho_vec z - y)_bot) into the Sp panel
            for (int r = tid >> 3; r < n8; r += kTileThreads / 8) {
              double v = 0.0;
              if (r < n) {
                const int sl = s8 ^ swz(r);              // the slot at this position of row r
                const double rho_s = C.rho[nb * 8 + sl];
                const int e1 = (nb * m8 + r) * 8 + s8, e2 = (nb * m8 + r + mp) * 8 + (sl ^ swz(r + mp)), en = (nb * n8 + r) * 8 + s8;
                const double w1 = rho_of((int)__ldg(P.ctype + r), rho_s) * zp[e1] - yp[e1];
                const double w2 = rho_of((int)__ldg(P.ctype + r + mp), rho_s) * zp[e2] - yp[e2];
                v = (S.sigma * cv[xi_at(en)] - qh[en]) + __ldg(K.adiag + r) * (w1 - w2);
              }
              Sp[(nb * n8 + r) * 8 + s8] = v;
            }
          } else if (PAIRED) {
            for (int r = tid >> 3; r < mp; r += kTileThreads / 8) {
              const int sl = s8 ^ swz(r);
              const double rho_s = C.rho[nb * 8 + sl];
              const int e1 = (nb * m8 + r) * 8 + s8, e2 = (nb * m8 + r + mp) * 8 + (sl ^ swz(r + mp));
              const double w1 = rho_of((int)__ldg(P.ctype + r), rho_s) * zp[e1] - yp[e1];
              const double w2 = rho_of((int)__ldg(P.ctype + r + mp), rho_s) * zp[e2] - yp[e2];
              cv[(nb * cvr + n8 + r) * 8 + s8] = w1 - w2;
            }
          } else {
            for (int r = tid >> 3; r < m8; r += kTileThreads / 8) {
              const int e = (nb * m8 + r) * 8 + s8;
              const double rho_s = C.rho[nb * 8 + (s8 ^ swz(r))];
              const int ct = r < m ? (int)__ldg(P.ctype + r) : 0;
              cv[(nb * cvr + n8 + r) * 8 + s8] = rho_of(ct, rho_s) * zp[e] - yp[e];
            }
          }
        }
      }
      if (tid < 32) {   // one lane per slot, combined with warp reductions
        int act = 0, ne = 0x7fffffff;
        if (tid < TB && C.inst[tid] >= 0) {
          act = 1;
          const int li = k - C.it0[tid];
          int nx = S.max_iter;
          if (check_every > 0) nx = min(nx, (li / check_every + 1) * check_every);
          if (adapt_every > 0) nx = min(nx, (li / adapt_every + 1) * adapt_every);
          ne = C.it0[tid] + nx;
        }
        act = __reduce_add_sync(0xffffffffu, act);
        ne = __reduce_min_sync(0xffffffffu, ne);
        if (tid == 0) { C.active = act; C.next_event = ne; C.refill_mask = 0; }
      }
      __syncthreads();
      initial = false;
#ifdef SMPC_TILE_PROFILE
      PFM(pf_pass, P_COUNT + 2)
      pf_t0 = clock64(); ++pf_nev;
      if (C.active == 0 && blockIdx.x == 0 && (tid == 0 || tid == (kTileWarps - 1) * 32)) {
        printf("tile profile: %d events, %d iterations %lld cycles\n  zero/flags %lld  w-rebuild/schedule %lld (gemm-tail %lld)\n", pf_nev, k, pf_it, pf_pass[P_COUNT], pf_pass[P_COUNT + 2], pf_pass[P_COUNT + 1]);
        for (int p = 0; p < P_COUNT; ++p) printf("  pass %d: pre %lld gemm+epilogue %lld\n", p, pf_pre[p], pf_pass[p]);
        printf("  warp %d iteration: gemm1+epilogue %lld barrier %lld gemm2+epilogue %lld barrier %lld\n", warp, pf_i[0], pf_i[1], pf_i[2], pf_i[3]);
        printf("  pass 8 pre: decide %lld store %lld tickets %lld load %lld badbounds %lld (rest = q̄)\n", pf_x[0], pf_x[1], pf_x[2], pf_x[3], pf_x[4]);
      }
#endif
      if (C.active == 0) break;
    }

    // ================= one ADMM iteration for the whole tile
    ++k;
    event = (k == C.next_event);
    if constexpr (XD) {
      const double2 *VTl = reinterpret_cast<const double2 *>(K.VTp) + lane, *Vl = reinterpret_cast<const double2 *>(K.Vp) + lane;
      // ---- GEMM 1: u = dinv .* (V' rhs)
      for (int rb = nrb0; rb < nrb1; rb += kRG) {
        double acc[kRG][NB][2];
        zero_acc(acc);
        gemm_seg<NB, kRG, kRing>(VTl, kpN, rb, min(kRG, nrb1 - rb), 0, kpN, Sp + bfrag, n8 * 8, acc);
#pragma unroll
        for (int r = 0; r < kRG; ++r)
          if (rb + r < nrb1) {
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
              const int pi = pidx(n8, nb, 8 * (rb + r) + g);
              const double2 dv = *reinterpret_cast<const double2 *>(dinv + pi);
              *reinterpret_cast<double2 *>(Tp + pi) = make_double2(acc[r][nb][0] * dv.x, acc[r][nb][1] * dv.y);
            }
          }
      }
      PFM(pf_i, 0)
      __syncthreads();
      PFM(pf_i, 1)
      // ---- GEMM 2: x~ = V u; then, per row i (the owner of x_i also owns constraint rows i and i + mp): x update, z~ = +-a_i x~_i,
      //      z / y updates (OSQP update_z / update_y), next right-hand side
      for (int rb = nrb0; rb < nrb1; rb += kRG) {
        double acc[kRG][NB][2];
        zero_acc(acc);
        gemm_seg<NB, kRG, kRing>(Vl, kpN, rb, min(kRG, nrb1 - rb), 0, kpN, Tp + bfrag, n8 * 8, acc);
#pragma unroll
        for (int r = 0; r < kRG; ++r)
          if (rb + r < nrb1) {
            const int row = 8 * (rb + r) + g;
            if (row < n) {
              const int ct1 = (int)__ldg(P.ctype + row), ct2 = (int)__ldg(P.ctype + row + mp);
              const double a = __ldg(K.adiag + row);
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) {
                const int s = nb * 8 + q2, pn = pidx(n8, nb, row), ci = pidx(cvr, nb, row);
                const double2 xo = *reinterpret_cast<const double2 *>(cv + ci), qv = *reinterpret_cast<const double2 *>(qh + pn);
                const double x0 = alpha * acc[r][nb][0] + oma * xo.x, x1 = alpha * acc[r][nb][1] + oma * xo.y;
                *reinterpret_cast<double2 *>(cv + ci) = make_double2(x0, x1);
                if (event) *reinterpret_cast<double2 *>(Dp + pn) = make_double2(x0 - xo.x, x1 - xo.y);
                const double zt0 = a * acc[r][nb][0], zt1 = a * acc[r][nb][1];
                double wsum2[2] = {0.0, 0.0};
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                  const int rr = row + half * mp, ct = half ? ct2 : ct1, pi = pidx(m8, nb, rr);
                  const double sg = half ? -1.0 : 1.0;
                  const double2 zo = *reinterpret_cast<const double2 *>(zp + pi), yo = *reinterpret_cast<const double2 *>(yp + pi);
                  const double2 lo = *reinterpret_cast<const double2 *>(lbp + pi), hi = *reinterpret_cast<const double2 *>(ubp + pi);
                  double rv0, rv1, ri0, ri1;
                  if (ct == 0) { rv0 = C.rho[s]; rv1 = C.rho[s + 1]; ri0 = C.rinv[s]; ri1 = C.rinv[s + 1]; }
                  else if (ct == 1) { rv0 = C.rho_eq[s]; rv1 = C.rho_eq[s + 1]; ri0 = C.rinv_eq[s]; ri1 = C.rinv_eq[s + 1]; }
                  else { rv0 = rv1 = kRhoMin; ri0 = ri1 = 1.0 / kRhoMin; }
                  const double zr0 = alpha * (sg * zt0) + oma * zo.x, zr1 = alpha * (sg * zt1) + oma * zo.y;
                  const double zn0 = fmin(fmax(zr0 + ri0 * yo.x, lo.x), hi.x), zn1 = fmin(fmax(zr1 + ri1 * yo.y, lo.y), hi.y);
                  const double d0 = rv0 * (zr0 - zn0), d1 = rv1 * (zr1 - zn1);
                  const double yn0 = yo.x + d0, yn1 = yo.y + d1;
                  *reinterpret_cast<double2 *>(zp + pi) = make_double2(zn0, zn1);
                  *reinterpret_cast<double2 *>(yp + pi) = make_double2(yn0, yn1);
                  if (event) *reinterpret_cast<double2 *>(cv + pidx(cvr, nb, n8 + rr)) = make_double2(d0, d1);   // delta_y for is_primal_infeasible
                  wsum2[0] += sg * (rv0 * zn0 - yn0); wsum2[1] += sg * (rv1 * zn1 - yn1);
                }
                *reinterpret_cast<double2 *>(Sp + pn) = make_double2((S.sigma * x0 - qv.x) + a * wsum2[0], (S.sigma * x1 - qv.y) + a * wsum2[1]);
              }
            }
          }
      }
      PFM(pf_i, 2)
      __syncthreads();
      PFM(pf_i, 3)
      continue;
    }
    const double2 *M1l = reinterpret_cast<const double2 *>(PAIRED ? K.M1p : K.M1) + lane;
    const double2 *Wl = reinterpret_cast<const double2 *>(PAIRED ? K.Wtop : K.Wp) + lane;
    const int kp1 = kpN + (PAIRED ? kpMp : kpM);      // k-pairs of GEMM 1
    // ---- GEMM 1: T = ([sigma G | W'] [xi; w] - q̂) .* dinv
    for (int rb = nrb0; rb < nrb1; rb += kRG) {
      double acc[kRG][NB][2];
      zero_acc(acc);
      gemm_seg<NB, kRG, kRing>(M1l, kp1, rb, min(kRG, nrb1 - rb), 0, kp1, cv + bfrag, cvr * 8, acc);
#pragma unroll
      for (int r = 0; r < kRG; ++r)
        if (rb + r < nrb1) {
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int pi = pidx(n8, nb, 8 * (rb + r) + g);
            const double2 qv = *reinterpret_cast<const double2 *>(qh + pi), dv = *reinterpret_cast<const double2 *>(dinv + pi);
            *reinterpret_cast<double2 *>(Tp + pi) = make_double2((acc[r][nb][0] - qv.x) * dv.x, (acc[r][nb][1] - qv.y) * dv.y);
          }
        }
    }
    __syncthreads();
    // ---- xi = alpha T + (1 - alpha) xi on the rows this warp produced (nobody reads xi in this phase);
    //      the iteration before an event also keeps delta_xi (is_dual_infeasible)
    for (int rb = nrb0; rb < nrb1; ++rb) {
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int row = 8 * rb + g, pi = pidx(n8, nb, row), ci = pidx(cvr, nb, row);
        const double2 t = *reinterpret_cast<const double2 *>(Tp + pi), xo = *reinterpret_cast<const double2 *>(cv + ci);
        const double x0 = alpha * t.x + oma * xo.x, x1 = alpha * t.y + oma * xo.y;
        *reinterpret_cast<double2 *>(cv + ci) = make_double2(x0, x1);
        if (event) *reinterpret_cast<double2 *>(Dp + pi) = make_double2(x0 - xo.x, x1 - xo.y);
      }
    }
    // ---- GEMM 2: z̃ = W T ; z, y updates (OSQP update_z / update_y) ; w = rho_vec z - y
    if (PAIRED) {
      // top half only: row p gives z̃_p, row p + mp gets -z̃_p; the panel row p of the w part carries w_p - w_{p+mp}
      for (int rb = prb0; rb < prb1; rb += kRG) {
        double acc[kRG][NB][2];
        zero_acc(acc);
        gemm_seg<NB, kRG, kRing>(Wl, kpN, rb, min(kRG, prb1 - rb), 0, kpN, Tp + bfrag, n8 * 8, acc);
#pragma unroll
        for (int r = 0; r < kRG; ++r)
          if (rb + r < prb1) {
            const int row = 8 * (rb + r) + g;
            if (row < mp) {
              const int ct1 = (int)__ldg(P.ctype + row), ct2 = (int)__ldg(P.ctype + row + mp);
#pragma unroll
              for (int nb = 0; nb < NB; ++nb) {
                const int s = nb * 8 + q2;
                double wsum2[2] = {0.0, 0.0};
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                  const int rr = row + half * mp, ct = half ? ct2 : ct1, pi = pidx(m8, nb, rr);
                  const double sg = half ? -1.0 : 1.0;
                  const double2 zo = *reinterpret_cast<const double2 *>(zp + pi), yo = *reinterpret_cast<const double2 *>(yp + pi);
                  const double2 lo = *reinterpret_cast<const double2 *>(lbp + pi), hi = *reinterpret_cast<const double2 *>(ubp + pi);
                  double rv0, rv1, ri0, ri1;
                  if (ct == 0) { rv0 = C.rho[s]; rv1 = C.rho[s + 1]; ri0 = C.rinv[s]; ri1 = C.rinv[s + 1]; }
                  else if (ct == 1) { rv0 = C.rho_eq[s]; rv1 = C.rho_eq[s + 1]; ri0 = C.rinv_eq[s]; ri1 = C.rinv_eq[s + 1]; }
                  else { rv0 = rv1 = kRhoMin; ri0 = ri1 = 1.0 / kRhoMin; }
                  const double zr0 = alpha * (sg * acc[r][nb][0]) + oma * zo.x, zr1 = alpha * (sg * acc[r][nb][1]) + oma * zo.y;
                  const double zn0 = fmin(fmax(zr0 + ri0 * yo.x, lo.x), hi.x), zn1 = fmin(fmax(zr1 + ri1 * yo.y, lo.y), hi.y);
                  const double d0 = rv0 * (zr0 - zn0), d1 = rv1 * (zr1 - zn1);
                  const double yn0 = yo.x + d0, yn1 = yo.y + d1;
                  *reinterpret_cast<double2 *>(zp + pi) = make_double2(zn0, zn1);
                  *reinterpret_cast<double2 *>(yp + pi) = make_double2(yn0, yn1);
                  // before an event the w panel carries delta_y of EVERY row (is_primal_infeasible); rebuilt after the event
                  if (event) *reinterpret_cast<double2 *>(cv + pidx(cvr, nb, n8 + rr)) = make_double2(d0, d1);
                  wsum2[0] += sg * (rv0 * zn0 - yn0); wsum2[1] += sg * (rv1 * zn1 - yn1);
                }
                if (!event) *reinterpret_cast<double2 *>(cv + pidx(cvr, nb, n8 + row)) = make_double2(wsum2[0], wsum2[1]);
              }
            }
          }
      }
      __syncthreads();
      continue;
    }
    for (int rb = mrb0; rb < mrb1; rb += kRG) {
      double acc[kRG][NB][2];
      zero_acc(acc);
      gemm_seg<NB, kRG, kRing>(Wl, kpN, rb, min(kRG, mrb1 - rb), 0, kpN, Tp + bfrag, n8 * 8, acc);
#pragma unroll
      for (int r = 0; r < kRG; ++r)
        if (rb + r < mrb1) {
          const int row = 8 * (rb + r) + g;
          const int ct = row < m ? (int)__ldg(P.ctype + row) : 0;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int pi = pidx(m8, nb, row), s = nb * 8 + q2;
            const double2 zo = *reinterpret_cast<const double2 *>(zp + pi), yo = *reinterpret_cast<const double2 *>(yp + pi);
            const double2 lo = *reinterpret_cast<const double2 *>(lbp + pi), hi = *reinterpret_cast<const double2 *>(ubp + pi);
            double rv0, rv1, ri0, ri1;
            if (ct == 0) { rv0 = C.rho[s]; rv1 = C.rho[s + 1]; ri0 = C.rinv[s]; ri1 = C.rinv[s + 1]; }
            else if (ct == 1) { rv0 = C.rho_eq[s]; rv1 = C.rho_eq[s + 1]; ri0 = C.rinv_eq[s]; ri1 = C.rinv_eq[s + 1]; }
            else { rv0 = rv1 = kRhoMin; ri0 = ri1 = 1.0 / kRhoMin; }
            const double zr0 = alpha * acc[r][nb][0] + oma * zo.x, zr1 = alpha * acc[r][nb][1] + oma * zo.y;
            const double zn0 = fmin(fmax(zr0 + ri0 * yo.x, lo.x), hi.x), zn1 = fmin(fmax(zr1 + ri1 * yo.y, lo.y), hi.y);
            const double d0 = rv0 * (zr0 - zn0), d1 = rv1 * (zr1 - zn1);
            const double yn0 = yo.x + d0, yn1 = yo.y + d1;
            *reinterpret_cast<double2 *>(zp + pi) = make_double2(zn0, zn1);
            *reinterpret_cast<double2 *>(yp + pi) = make_double2(yn0, yn1);
            // before an event the w panel carries delta_y (is_primal_infeasible); w is rebuilt after the event
            *reinterpret_cast<double2 *>(cv + pidx(cvr, nb, n8 + row)) = event ? make_double2(d0, d1) : make_double2(rv0 * zn0 - yn0, rv1 * zn1 - yn1);
          }
        }
    }
    __syncthreads();
  }
}

bool tile_kernel_supports(int n, int m) { return tile_kernel_nb(n, m) > 0; }

size_t tile_smem_bytes(int n, int m, int nb) {
  const size_t n8 = (n + 7) & ~7, m8 = (m + 7) & ~7, TB = 8 * nb;
  const size_t ctl = nb == 1 ? sizeof(TileCtl<8, 16>) : sizeof(TileCtl<16, 16>);   // (the larger of the 8- and 16-warp control blocks)
  return (6 * n8 + 5 * m8) * TB * sizeof(double) + ctl;
}

int tile_kernel_nb(int n, int m) {
  if (n < 1 || m < 0) return 0;
  const size_t cap = 227 * 1024;
  if (tile_smem_bytes(n, m, 2) <= cap) return 2;
  if (tile_smem_bytes(n, m, 1) <= cap) return 1;
  return 0;
}

cudaError_t launch_admm_shared_tile(const TilePackDev &K, const SharedPlanDev &P, const BatchDev &Bt, const SettingsDev &S,
                                    int *queue, int nb, int num_sms, cudaStream_t stream) {
  if (nb == 0) nb = tile_kernel_nb(P.n, P.m);
  if (nb != 1 && nb != 2) return cudaErrorInvalidValue;
  const bool paired = K.mp > 0;
  const size_t cap = 227 * 1024;
  // Two independent 8-warp tiles per SM (the kernel's CTAS = 2 variant: 128 registers per thread) when two tiles fit shared memory,
  // if need be with one 8-slot block each instead of two: their barriers and events are not synchronised, so one tile's check /
  // store / refill overlaps the other's GEMMs (config 5, N = 100: 7.40 -> 6.88 ms per step).  SMPC_TILE_CTAS=1 switches it off.
  static const int want_ctas = [] { const char *e = getenv("SMPC_TILE_CTAS"); return e ? atoi(e) : 2; }();
  int ctas = 1;
  if (want_ctas == 2 && paired && !K.xd) {
    if (2 * (tile_smem_bytes(P.n, P.m, nb) + 1024) <= cap) ctas = 2;
    else if (nb == 2 && 2 * (tile_smem_bytes(P.n, P.m, 1) + 1024) <= cap) { nb = 1; ctas = 2; }
  }
  const size_t smem = tile_smem_bytes(P.n, P.m, nb);
  if (smem > cap) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(queue, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  const int TB = 8 * nb;
  int grid = (Bt.B + TB - 1) / TB;
  // one CTA per SM when the tile fills shared memory; small problems let several CTAs share an SM
  int per_sm = (int)(cap / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  if (ctas == 2) per_sm = 2;
  if (grid > num_sms * per_sm) grid = num_sms * per_sm;
  // tiles that own an SM alone run 16 warps (4 per sub-partition cover each other's LDS / L2 / barrier stalls: +11 % at N = 100,
  // +18 % on the quadrotor); smaller tiles keep 8 warps with 4 row-blocks per warp
  int warps = (per_sm == 1 || K.n8 >= 48) ? 16 : 8;   // (with >= 6 row-blocks of n the 16 warps all have work; measured N = 30: 8 warps, N = 50: 16)
  if (const char *env = getenv("SMPC_TILE_WARPS")) warps = atoi(env) == 16 ? 16 : (atoi(env) == 8 ? 8 : warps);   // development knob
  if (ctas == 2) warps = 8;
  auto go = [&](auto kernel) -> cudaError_t {
    cudaError_t e2 = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e2 != cudaSuccess) return e2;
    kernel<<<grid, warps * 32, smem, stream>>>(K, P, Bt, S, queue);
    return cudaGetLastError();
  };
  if (ctas == 2) return nb == 1 ? go(admm_shared_tile_kernel<1, true, 8, false, 2>) : go(admm_shared_tile_kernel<2, true, 8, false, 2>);
  if (K.xd && paired) {
    if (warps == 16) return nb == 1 ? go(admm_shared_tile_kernel<1, true, 16, true>) : go(admm_shared_tile_kernel<2, true, 16, true>);
    return nb == 1 ? go(admm_shared_tile_kernel<1, true, 8, true>) : go(admm_shared_tile_kernel<2, true, 8, true>);
  }
  if (warps == 16) {
    if (nb == 1) return paired ? go(admm_shared_tile_kernel<1, true, 16>) : go(admm_shared_tile_kernel<1, false, 16>);
    return paired ? go(admm_shared_tile_kernel<2, true, 16>) : go(admm_shared_tile_kernel<2, false, 16>);
  }
  if (nb == 1) return paired ? go(admm_shared_tile_kernel<1, true, 8>) : go(admm_shared_tile_kernel<1, false, 8>);
  return paired ? go(admm_shared_tile_kernel<2, true, 8>) : go(admm_shared_tile_kernel<2, false, 8>);
}

}  // namespace smpc
