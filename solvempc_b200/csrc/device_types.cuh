// device_types.cuh -- POD views handed to the kernels (all pointers are device pointers).
#pragma once
#include <cuda_runtime.h>

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
};

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
