/*
 * oracle/mpc_assembly.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's condensed-QP builders
 * (/root/reference/src/ModelPredictiveControlAPI.cpp, cited per function) with
 * the compile-time dimensions mpcWindow / N_S (include/ModelPredictiveControlAPI.h:26-32)
 * lifted to run-time N / nx.  N_C = N_O = 1 as in the reference.
 * All matrices row-major doubles.  Validated against the reference's own
 * compiled .cpp (oracle/_ref, tests/golden/assembly_ref.json).
 *
 * Reference quirks reproduced on purpose:
 *  - S holds K only in its first `n_state_rows` rows (literal 10 at cpp:185),
 *    the remaining rows are zero;
 *  - Su's strict upper triangle is zero (never written, cpp:197-204);
 *  - Fu uses diag(LL' * Rbar') = R * 1, not LL' * Rbar * 1 (cpp:305).
 */
#include "mpc_assembly.h"

#include <float.h>
#include <stdlib.h>
#include <string.h>

void orc_mpc_build(int N, int nx, const double *Ad, const double *Bd, const double *Cd,
                   const double *K, double Q, double R, double RD, int n_state_rows,
                   double u_limit, double *H, double *Gbar, double *Fx, double *Fu,
                   double *Fr, double *Sbar, double *Ku, double *W0, double *Sx,
                   double *Su, double *CAB) {
  double *Ak = (double *)calloc((size_t)nx * nx, sizeof(double));   /* Ad^i */
  double *An = (double *)calloc((size_t)nx * nx, sizeof(double));
  double *row = (double *)calloc(nx, sizeof(double));
  for (int i = 0; i < nx; i++) Ak[i * nx + i] = 1.0;
  /* setTransformations, cpp:187-190: Sx[i] = Cd*Ad^(i+1); CAB[i] = Cd*Ad^i*Bd */
  for (int i = 0; i < N; i++) {
    for (int c = 0; c < nx; c++) { double s = 0; for (int k = 0; k < nx; k++) s += Cd[k] * Ak[k * nx + c]; row[c] = s; } /* Cd*Ad^i */
    { double s = 0; for (int k = 0; k < nx; k++) s += row[k] * Bd[k]; CAB[i] = s; }
    for (int r = 0; r < nx; r++) for (int c = 0; c < nx; c++) { double s = 0; for (int k = 0; k < nx; k++) s += Ak[r * nx + k] * Ad[k * nx + c]; An[r * nx + c] = s; }
    memcpy(Ak, An, sizeof(double) * nx * nx);
    for (int c = 0; c < nx; c++) { double s = 0; for (int k = 0; k < nx; k++) s += Cd[k] * Ak[k * nx + c]; Sx[i * nx + c] = s; }
  }
  /* cpp:197-201: Su(i,j) = sum_{k<=i-j} CAB[k], j<=i; strict upper = 0 */
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) {
    double s = 0; if (j <= i) for (int k = 0; k <= i - j; k++) s += CAB[k];
    Su[i * N + j] = s;
  }
  /* cpp:185,208: S rows [0,n_state_rows) = K, rest 0; Sbar = [S; -S] */
  int ns = n_state_rows < N ? n_state_rows : N;
  for (int i = 0; i < N; i++) for (int c = 0; c < nx; c++) {
    double v = i < ns ? K[c] : 0.0;
    Sbar[i * nx + c] = v; Sbar[(N + i) * nx + c] = -v;
  }
  /* setH, cpp:250-251: H = sym(2*(LL'*Rbar*LL + RbarD + Su'*Qbar*Su)); (LL'LL)(i,j) = N - max(i,j) */
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) {
    double s = 0; for (int k = 0; k < N; k++) s += Su[k * N + i] * Su[k * N + j];
    int mx = i > j ? i : j;
    H[i * N + j] = 2.0 * (R * (double)(N - mx) + (i == j ? RD : 0.0) + Q * s);
  }
  for (int i = 0; i < N; i++) for (int j = i + 1; j < N; j++) { double v = (H[i * N + j] + H[j * N + i]) / 2.0; H[i * N + j] = v; H[j * N + i] = v; }
  /* setFVars, cpp:305-307 */
  for (int i = 0; i < N; i++) {
    double s = 0; for (int k = 0; k < N; k++) s += Su[k * N + 0] * Su[k * N + i];
    Fu[i] = 2.0 * (R + Q * s);
    for (int j = 0; j < N; j++) Fr[i * N + j] = -2.0 * Q * Su[j * N + i];
    for (int c = 0; c < nx; c++) { double t = 0; for (int k = 0; k < N; k++) t += Su[k * N + i] * Sx[k * nx + c]; Fx[i * nx + c] = 2.0 * Q * t; }
  }
  /* setLinearConstraints, cpp:332-335: Gbar = [K(0)*LL ; -K(0)*LL] */
  for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) {
    double v = j <= i ? K[0] : 0.0;
    Gbar[i * N + j] = v; Gbar[(N + i) * N + j] = -v;
  }
  /* setUpperBound, cpp:364-368: Ku = [-K(0)*1 ; K(0)*1]; W0 = 255*1 */
  for (int i = 0; i < N; i++) { Ku[i] = -K[0]; Ku[N + i] = K[0]; W0[i] = u_limit; W0[N + i] = u_limit; }
  free(Ak); free(An); free(row);
}

/* setF (cpp:374) and the upper bound sent to the solver (cpp:99); lb = -DBL_MAX (cpp:42) */
void orc_mpc_step_vectors(int N, int nx, const double *Fx, const double *Fu, const double *Fr,
                          const double *Sbar, const double *Ku, const double *W0,
                          const double *X, double U, const double *ref, double *f, double *ub) {
  for (int i = 0; i < N; i++) {
    double s = 0; for (int c = 0; c < nx; c++) s += Fx[i * nx + c] * X[c];
    s += Fu[i] * U;
    double r = 0; for (int j = 0; j < N; j++) r += Fr[i * N + j] * ref[j];
    f[i] = s + r;
  }
  for (int i = 0; i < 2 * N; i++) {
    double s = 0; for (int c = 0; c < nx; c++) s += Sbar[i * nx + c] * X[c];
    ub[i] = (W0[i] + s) + Ku[i] * U;
  }
}

double orc_mpc_lower_bound(void) { return -DBL_MAX; }
