// device_types.cuh -- POD views handed to the kernels (all pointers are device pointers).
#pragma once
#include <cuda_runtime.h>

// -DSMPC_DEBUG_BOUNDS (make NVEXTRA=-DSMPC_DEBUG_BOUNDS): in-kernel checks of the work-queue / ticket protocols and of every
// index taken from them (compute-sanitizer is not available on the GPU pool); a failed check prints its site and traps, which
// the host sees as a launch failure.  tools/run_debug_bounds.sh runs the GPU tests against such a build.
#ifdef SMPC_DEBUG_BOUNDS
#include <cstdio>
#define SMPC_DBG(cond, what) do { if (!(cond)) { printf("SMPC_DEBUG_BOUNDS: %s failed (%s:%d)\n", what, __FILE__, __LINE__); __trap(); } } while (0)
#else
#define SMPC_DBG(cond, what) do { } while (0)
#endif

namespace smpc {

constexpr double kRhoMin = 1e-6, kRhoMax = 1e6, kRhoEqOverIneq = 1e3;
constexpr double kInfty = 1e30, kMinScaling = 1e-4, kDivTol = 1e-10, kRhoTolRow = 1e-4;

struct SettingsDev {
  double rho0, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf, rho_tol;
  int max_iter, check_every, adaptive_rho, rho_interval, warm_start, scaled_termination;
};

// shared-factor plan (see plan.hpp); matrices are stored so that lane = output row reads coalesced
struct SharedPlanDev {
  int n, m;
  double c, cinv;
  const double *SG;     // n*n   sigma*V'V (symmetric)
  const double *W;      // m*n   A̅V, row-major:  t-phase reads W[r*n+i]
  const double *WT;     // n*m   (A̅V)':          z-phase reads WT[k*m+r]
  const double *V;      // n*n   row-major:       qhat_i = sum_k V[k*n+i] qbar_k
  const double *VT;     // n*n                    xbar_i = sum_k VT[k*n+i] xi_k
  const double *PVT;    // n*n                    (P̄x)_i = sum_k PVT[k*n+i] xi_k
  const double *VinvT;  // n*n                    xi_i = sum_k VinvT[k*n+i] xbar_k
  const double *Abar;   // m*n   row-major:       (A̅'y)_i = sum_r Abar[r*n+i] y_r
  const double *Pbar;   // n*n   full symmetric (polish only)
  const double *lam, *D, *Dinv, *E, *Einv;
  const double *l0, *u0;        // UNSCALED setup bounds (used when the batch has none of its own)
  const signed char *ctype;     // m
  double dx_bound;              // ||diag(D) V||_inf (||V||_inf with scaled_termination): ||delta_x||_inf <= dx_bound ||delta_xi||_inf
};

// zero-padded, k-major packs of the plan for the register-resident small-QP kernel (n<=16, m<=32)
struct SmallPackDev {
  const double *M1T;   // [48][16]  M1T[k][i] = [sigma*G | W'](i,k)
  const double *WT;    // [16][32]  WT[k][r]  = W(r,k)
  const double *VT;    // [16][16]  VT[k][i]  = V(i,k)
  const double *PVT;   // [16][16]  PVT[k][i] = (P̄V)(i,k)
  const double *Ab;    // [32][16]  A̅ row-major
  const double *V;     // [16][16]  V row-major
  const double *lam, *D, *Dinv;   // [16]
  const double *E, *Einv;         // [32]
  const int *ctype;               // [32]
  int mp;                         // > 0: rows are [G; -G] with mp = m / 2 pairs (row mp + i = -row i), else 0
  // packs of the fused one-phase kernel (admm_shared_small_fused.cu; paired plans with mp <= 16 only, else NULL).  Lane (a, b) =
  // (lane >> 2, lane & 3) of a warp owns entry idx = 2a + (b & 1) of xi (b < 2, "xi-lane") or pair idx (b >= 2, "pair-lane").
  const double *M1x;   // [16][32]  M1x[k][small_pos32(e)] = [sigma*G | Wtop'](k, e): column order of the kernel's state s = [xi; wd]
  const double *WTt;   // [16][16]  WTt[k][i] = Wtop(i, k)
  const double *C1;    // [8][32][2] 4 x 4 blocks of [V; Wtop] (32 x 16) per lane, chunk-major (layout: upload_small_pack)
  const double *C2;    // [8][32][2] 4 x 4 blocks of [P̄V; A̅top'] (32 x 16)
  const double *cst;   // [32][4]   per lane: xi-lane {D, 1/D, lambda, 0} of entry idx; pair-lane {E, 1/E of row idx, E, 1/E of row mp + idx}
};
// storage order of a 32-vector read as four 16-byte chunks per lane (lane b reads entries 8b .. 8b+7: the four chunks one
// LDS.128 touches are contiguous, no bank conflict), and of a 16-vector read as two chunks per lane (entries 4b .. 4b+3)
__host__ __device__ inline int small_pos32(int e) { return ((((e & 7) >> 1) * 4 + (e >> 3)) * 2) + (e & 1); }
__host__ __device__ inline int small_pos16(int e) { return ((((e & 3) >> 1) * 4 + (e >> 2)) * 2) + (e & 1); }
// local row l (0..3) of lane (a, b) in the 2-D block layout is the row finally owned by lane (a, b ^ small_rowmix(l))
__host__ __device__ inline int small_rowmix(int l) { return ((l & 1) << 1) | (l >> 1); }

// DMMA A-fragment packs of the plan for the tile kernel (admm_shared_tile.cu).  An operator Op (rows x K, both
// zero-padded to multiples of 8) is stored as [row-block][k-pair][lane][2]:
//   pack[((rb*(K8/8) + kp)*32 + lane)*2 + j] = Op[8 rb + lane/4][8 kp + 4 j + lane%4]
// i.e. one 16-byte load per lane feeds two mma.sync.m8n8k4.f64 A operands.
struct TilePackDev {
  int n8, m8;          // n, m rounded up to multiples of 8
  const double *M1;    // [sigma*G | W']  n8 x (n8 + m8)
  const double *Wp;    // W = A̅V          m8 x n8
  const double *VTp;   // V'              n8 x n8   (q̂ = V' q̄)
  const double *Vp;    // V               n8 x n8   (x̄ = V xi)
  const double *PVp;   // P̄V              n8 x n8
  const double *ATp;   // A̅'              n8 x m8
  // Paired rows (mp > 0): the reference writes its two-sided limit as the row sets [G; -G] (cpp:335), so row r + mp of
  // A̅ -- and of W -- is exactly the negative of row r (mp = m / 2).  Then W'w = Wtop'(w_top - w_bot) and the bottom half of
  // z̃ = W t is -z̃_top: the iteration GEMMs run on Wtop only (K = n8 + mp8 and mp8 rows), 1.67 x fewer flops at m = 2n.
  int mp, mp8;
  const double *M1p;   // [sigma*G | Wtop']  n8 x (n8 + mp8)
  const double *Wtop;  // first mp rows of W  mp8 x n8
  const double *ATtop; // first mp columns of A̅' (= G')  n8 x mp8: A̅'y = G'(y_top - y_bot) in the checks
  // x-space variant (xd != 0): paired rows whose top block is DIAGONAL after scaling, A̅ = [diag(adiag); -diag(adiag)] (mp == n):
  // the iteration multiplies by V' and V only (2 n^2 MACs) and needs P̄ for the checks
  int xd;
  const double *Pp;     // P̄  n8 x n8 (fragment pack)
  const double *adiag;  // n8 (zero padded)
  double dmax;          // max_i D_i (unscaled termination) or 1: ||delta_x||_inf <= dmax ||delta_x̄||_inf
};

// per-instance regime: every QP has its own (scaled) P̄_i, A̅_i and scaling
struct InstanceDataDev {
  int n, m, B;
  double *P;            // [B][n][n]  P̄ (full symmetric) after ruiz_instance_kernel
  double *A;            // [B][m][n]  A̅
  double *D, *E, *c;    // [B][n], [B][m], [B]
  const double *l0, *u0;   // shared UNSCALED setup bounds (m)
  // prepared at create time by the register-operator kernel (NULL when it does not apply): M(rho) = S0 + rho T for the
  // row classes of the setup bounds, packed lower triangles [B][n(n+1)/2], and the rows of M(rho_prepared)^-1, k-major [B][32][n]
  double *S0, *T, *Minv0;
  double rho_prepared;
  int paired;           // 1: rows r and r + m/2 of every scaled A̅_i are exact negatives ([G; -G], cpp:335), checked at create time
  // two-warp kernel for paired instances (admm_instance_pair.cu), NULL when it does not apply: per-instance pack
  // [G' | M(pack_rho)^-1 | K = G M^-1 | S0 | T] (instance_pair_pack_doubles(n) doubles each), the rho each pack was last factored
  // with, and the ticket queue (2 ints)
  double *pack, *pack_rho;
  int *queue;
};

// what the polish kernel reads (polish.cu): scaled data and scaling of either regime; strides are 0 when the batch shares them
struct PolishDataDev {
  int n, m;
  const double *Pbar, *Abar;    // n*n full symmetric, m*n row-major
  const double *D, *E;          // n, m
  size_t strideP, strideA, strideD, strideE;
  double c;                     // cost scaling when shared
  const double *c_inst;         // [B] per-instance cost scaling, or NULL
  const double *l0, *u0;        // UNSCALED setup bounds (m)
  const double *VinvT;          // shared-factor kernels that keep xi = V^-1 x̄ as their state; NULL: the state is x̄ itself
};

// per-instance data and state, [B][len] contiguous
struct BatchDev {
  int B;
  int fresh;            // 1: ignore stored iterates and rho (x = z = y = 0, rho = rho0)
  const double *q;      // [B][n] unscaled gradient (NULL = 0)
  const double *l, *u;  // [B][m] unscaled bounds (NULL = plan.l0 / plan.u0)
  double *xi;           // [B][n] warm-start state in plan coordinates
  double *z, *y;        // [B][m] scaled z, y
  double *rho;          // [B]
  double *x_out, *y_out;                 // [B][n], [B][m] unscaled solution
  int *status, *iter, *rho_updates;      // [B]
  double *obj, *pri_res, *dua_res;       // [B]
  double *u_apply;      // [B] or NULL: MPC layer's U, incremented by x[0] of every instance that ends SOLVED (cpp:105);
                        // honoured by the small-QP kernels only (the API launches mpc_apply_control_kernel otherwise)
  double *u_export;     // [B] or NULL: where the caller wants U after the step (device or pinned host, smpc_mpc_bind_results) and
  int *status_export;   // [B] or NULL: the statuses; written by admm_shared_small_kernel as each instance ends
};

}  // namespace smpc
