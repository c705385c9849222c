// plan.hpp -- host-side "plan" of the shared-factor regime.
//
// When every QP of the batch shares P and A (reference: one plant, one Hessian H and one
// constraint matrix Gbar built once in the constructor, src/ModelPredictiveControlAPI.cpp:27-39,
// handed to the solver at cpp:57-59), everything OSQP computes at osqp_setup is shared too:
// the Ruiz equilibration (D, E, c) and the KKT factor.  OSQP refactors the KKT matrix whenever
// an instance adapts its rho; to keep ONE set of shared matrices for ALL per-instance rho values
// the plan diagonalises the pencil (S, T):
//     M(rho) = P̄ + sigma I + A̅' diag(rho_vec) A̅ = S + rho T,
//     S = P̄ + sigma I + rho_min A̅_f' A̅_f  (free rows),  T = A̅' diag(kappa) A̅  (kappa = 1 | 1e3 | 0)
//     V' S V = I,  V' T V = diag(lambda)   =>   M(rho)^-1 = V diag(1/(1+rho*lambda)) V'.
// In the coordinates x̄ = V xi the KKT solve of every ADMM iteration is a diagonal scaling and the
// iteration needs only the shared dense operators sigma*G = sigma V'V, W = A̅ V (see admm_shared.cu).
// This is the same arithmetic OSQP performs (exact for every rho), not an approximation.
#pragma once
#include <string>
#include <vector>

#include "../../include/solvempc_b200.h"

namespace smpc {

struct SharedPlan {
  int n = 0, m = 0;
  double c = 1.0, cinv = 1.0;
  std::vector<double> D, Dinv, E, Einv;     // Ruiz scaling
  std::vector<double> Pbar, Abar;           // scaled data, row-major n*n, m*n
  std::vector<double> l0bar, u0bar;         // scaled setup bounds
  std::vector<signed char> ctype;           // -1 free, 0 inequality, 1 equality (OSQP constr_type)
  std::vector<double> lam;                  // generalised eigenvalues (n)
  std::vector<double> V, VT;                // V row-major, and its transpose
  std::vector<double> SG;                   // sigma * V'V           (n*n, symmetric)
  std::vector<double> W, WT;                // A̅ V (m*n) and transpose (n*m)
  std::vector<double> PVT;                  // (P̄ V)' (n*n)
  std::vector<double> VinvT;                // (V' S)' = S V (n*n): xi = Vinv x̄
  int pairs = 0;                            // rows r, r+m/2 with A[r+m/2] == -A[r] (diagnostic)
  std::vector<double> xdiag;                // non-empty: pairs == n and the top block of A̅ is diag(xdiag) (x-space tile iteration)
};

// OSQP scale_data (modified Ruiz + cost scaling) on (P, A, q): returns D, E, c and scaled P̄, A̅.
void ruiz_scale(int n, int m, int iters, std::vector<double> &P, std::vector<double> &A, std::vector<double> &q,
                std::vector<double> &D, std::vector<double> &E, double &c);

// Builds the plan; returns SMPC_OK or an error code with `err` filled.
int build_shared_plan(int n, int m, const double *P_upper_rowmajor, const double *A_rowmajor,
                      const double *q0, const double *l0, const double *u0, const smpc_settings &st, SharedPlan &plan,
                      std::string &err);

// Symmetric eigen-decomposition C = Q diag(w) Q' by cyclic Jacobi (C row-major n*n is destroyed).
void jacobi_eigh(int n, std::vector<double> &C, std::vector<double> &w, std::vector<double> &Q);

}  // namespace smpc
