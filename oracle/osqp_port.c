/*
 * oracle/osqp_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see osqp_port.h).
 *
 * Plain-C restatement of the OSQP 0.6.x solver that the reference calls through
 * osqp-eigen at src/ModelPredictiveControlAPI.cpp:51-64 (setup) and :96-105
 * (update + solve).  Section numbers below refer to SURVEY.md section 3.4 and
 * to the OSQP 0.6.x source files whose published algorithm is restated
 * (third-party, not under /root/reference, version unpinned by the reference).
 *
 * PARITY UNPINNED against real osqp-eigen; pinned by exact KKT known answers.
 *
 * Linear system: OSQP factors the quasi-definite KKT matrix
 *     [ P+sigma*I   A' ; A   -diag(1/rho_vec) ]
 * with QDLDL after an AMD permutation.  The (2,2) block is diagonal, so the
 * fill-minimising elimination order is "constraint rows first"; LDL' in that
 * order is exactly block elimination: D_nu = -1/rho_vec, L21 = -A'*diag(rho_vec),
 * Schur complement S = P + sigma*I + A'*diag(rho_vec)*A factored as L*D*L'.
 * This file performs the LDL' solve in that order (no iterative refinement,
 * as OSQP with QDLDL does none).
 */
#define _POSIX_C_SOURCE 200809L
#include "osqp_port.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* osqp/include/constants.h */
#define RHO_MIN 1e-06
#define RHO_MAX 1e06
#define RHO_EQ_OVER_RHO_INEQ 1e03
#define RHO_TOL 1e-04
#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-04
#define MAX_SCALING 1e+04
#define OSQP_DIVISION_TOL 1e-10

struct orc_solver {
  int n, m;
  orc_settings st;
  double rho0;
  /* scaled data */
  double *P;    /* n*n full symmetric, row-major */
  double *A;    /* m*n row-major */
  double *q, *l, *u;
  /* CSR of A and of A' (skip structural zeros in mat-vecs) */
  int *a_rp, *a_ci; double *a_v;
  int *at_rp, *at_ci; double *at_v;
  /* scaling */
  double *D, *Dinv, *E, *Einv; double c, cinv;
  /* rho vectors */
  double *rho_vec, *rho_inv_vec; int *constr_type;
  /* factor of S = P + sigma I + A' R A :  S = L D L' (unit lower L) */
  double *L, *Dl;       /* n*n, n */
  double *L0, *Dl0;     /* factor at rho0 (orc_reset) */
  double *rho_vec0, *rho_inv_vec0; int *constr_type0;
  /* iterates */
  double *x, *z, *y, *x_prev, *z_prev, *xt, *zt; /* xt,zt = x_tilde, z_tilde */
  double *delta_x, *delta_y;
  double *Ax, *Px, *Aty, *Adelta_x, *Atdelta_y, *Pdelta_x;
  double *tn, *tm; /* temporaries */
  /* info */
  int status_val, iter, rho_updates, status_polish;
  double obj_val, pri_res, dua_res, rho_estimate;
  /* solution (unscaled) */
  double *sol_x, *sol_y;
};

void orc_default_settings(orc_settings *s) {
  s->rho = 0.1; s->sigma = 1e-6; s->alpha = 1.6;
  s->eps_abs = 1e-3; s->eps_rel = 1e-3;
  s->eps_prim_inf = 1e-4; s->eps_dual_inf = 1e-4;
  s->adaptive_rho_tolerance = 5.0;
  s->max_iter = 4000; s->check_termination = 25; s->scaling = 10;
  s->adaptive_rho = 1; s->adaptive_rho_interval = 25;
  s->warm_start = 1; s->scaled_termination = 0;
  s->polish = 0; s->polish_refine_iter = 3; s->delta = 1e-6;
}

static double *dalloc(size_t k) { return (double *)calloc(k ? k : 1, sizeof(double)); }
static double c_max(double a, double b) { return a > b ? a : b; }
static double c_min(double a, double b) { return a < b ? a : b; }
static double norm_inf(const double *v, int k) {
  double r = 0; for (int i = 0; i < k; i++) { double a = fabs(v[i]); if (a > r) r = a; } return r;
}
static double scaled_norm_inf(const double *s, const double *v, int k) {
  double r = 0; for (int i = 0; i < k; i++) { double a = fabs(s[i] * v[i]); if (a > r) r = a; } return r;
}

/* y = A x  (CSR) */
static void csr_mv(int rows, const int *rp, const int *ci, const double *v, const double *x, double *y) {
  for (int i = 0; i < rows; i++) {
    double s = 0; for (int k = rp[i]; k < rp[i + 1]; k++) s += v[k] * x[ci[k]]; y[i] = s;
  }
}
static void dense_sym_mv(int n, const double *P, const double *x, double *y) {
  for (int i = 0; i < n; i++) { double s = 0; const double *r = P + (size_t)i * n; for (int j = 0; j < n; j++) s += r[j] * x[j]; y[i] = s; }
}

static void build_csr(orc_solver *w) {
  int n = w->n, m = w->m, nnz = 0;
  for (int i = 0; i < m * n; i++) if (w->A[i] != 0.0) nnz++;
  free(w->a_rp); free(w->a_ci); free(w->a_v); free(w->at_rp); free(w->at_ci); free(w->at_v);
  w->a_rp = (int *)calloc(m + 1, sizeof(int)); w->a_ci = (int *)calloc(nnz ? nnz : 1, sizeof(int)); w->a_v = dalloc(nnz);
  w->at_rp = (int *)calloc(n + 1, sizeof(int)); w->at_ci = (int *)calloc(nnz ? nnz : 1, sizeof(int)); w->at_v = dalloc(nnz);
  int k = 0;
  for (int i = 0; i < m; i++) { w->a_rp[i] = k; for (int j = 0; j < n; j++) if (w->A[(size_t)i * n + j] != 0.0) { w->a_ci[k] = j; w->a_v[k] = w->A[(size_t)i * n + j]; k++; } }
  w->a_rp[m] = k; k = 0;
  for (int j = 0; j < n; j++) { w->at_rp[j] = k; for (int i = 0; i < m; i++) if (w->A[(size_t)i * n + j] != 0.0) { w->at_ci[k] = i; w->at_v[k] = w->A[(size_t)i * n + j]; k++; } }
  w->at_rp[n] = k;
}

/* osqp/src/scaling.c: limit_scaling */
static void limit_scaling(double *v, int k) {
  for (int i = 0; i < k; i++) { v[i] = v[i] < MIN_SCALING ? 1.0 : v[i]; v[i] = v[i] > MAX_SCALING ? MAX_SCALING : v[i]; }
}

/* osqp/src/scaling.c: scale_data (modified Ruiz equilibration + cost scaling) */
static void scale_data(orc_solver *w) {
  int n = w->n, m = w->m;
  double *Dt = w->tn, *Et = w->tm;
  w->c = 1.0;
  for (int i = 0; i < n; i++) w->D[i] = w->Dinv[i] = 1.0;
  for (int i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1.0;
  for (int it = 0; it < w->st.scaling; it++) {
    /* compute_inf_norm_cols_KKT: D_temp = max(col norms of P, col norms of A); E_temp = row norms of A */
    for (int j = 0; j < n; j++) {
      double r = 0;
      for (int i = 0; i < n; i++) r = c_max(r, fabs(w->P[(size_t)i * n + j]));
      for (int i = 0; i < m; i++) r = c_max(r, fabs(w->A[(size_t)i * n + j]));
      Dt[j] = r;
    }
    for (int i = 0; i < m; i++) { double r = 0; for (int j = 0; j < n; j++) r = c_max(r, fabs(w->A[(size_t)i * n + j])); Et[i] = r; }
    limit_scaling(Dt, n); limit_scaling(Et, m);
    for (int j = 0; j < n; j++) Dt[j] = 1.0 / sqrt(Dt[j]);
    for (int i = 0; i < m; i++) Et[i] = 1.0 / sqrt(Et[i]);
    /* P <- D P D ; A <- E A D ; q <- D q */
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) w->P[(size_t)i * n + j] = (Dt[i] * w->P[(size_t)i * n + j]) * Dt[j];
    for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) w->A[(size_t)i * n + j] = (Et[i] * w->A[(size_t)i * n + j]) * Dt[j];
    for (int j = 0; j < n; j++) w->q[j] = Dt[j] * w->q[j];
    for (int j = 0; j < n; j++) w->D[j] *= Dt[j];
    for (int i = 0; i < m; i++) w->E[i] *= Et[i];
    /* cost normalisation */
    double mean = 0;
    for (int j = 0; j < n; j++) { double r = 0; for (int i = 0; i < n; i++) r = c_max(r, fabs(w->P[(size_t)i * n + j])); mean += r; }
    mean /= n;
    double nq = norm_inf(w->q, n); limit_scaling(&nq, 1);
    double ct = c_max(mean, nq); limit_scaling(&ct, 1);
    ct = 1.0 / ct;
    for (int i = 0; i < n * n; i++) w->P[i] *= ct;
    for (int j = 0; j < n; j++) w->q[j] *= ct;
    w->c *= ct;
  }
  w->cinv = 1.0 / w->c;
  for (int j = 0; j < n; j++) w->Dinv[j] = 1.0 / w->D[j];
  for (int i = 0; i < m; i++) w->Einv[i] = 1.0 / w->E[i];
  for (int i = 0; i < m; i++) { w->l[i] = w->E[i] * w->l[i]; w->u[i] = w->E[i] * w->u[i]; }
}

/* osqp/src/auxil.c: set_rho_vec */
static void set_rho_vec(orc_solver *w) {
  w->st.rho = c_min(c_max(w->st.rho, RHO_MIN), RHO_MAX);
  for (int i = 0; i < w->m; i++) {
    if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) {
      w->constr_type[i] = -1; w->rho_vec[i] = RHO_MIN;
    } else if (w->u[i] - w->l[i] < RHO_TOL) {
      w->constr_type[i] = 1; w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->st.rho;
    } else {
      w->constr_type[i] = 0; w->rho_vec[i] = w->st.rho;
    }
    w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
  }
}

/* numeric LDL' of S = P + sigma I + A' diag(rho_vec) A (see file header) */
static int factorize(orc_solver *w) {
  int n = w->n, m = w->m;
  double *S = w->L;
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) S[(size_t)i * n + j] = w->P[(size_t)i * n + j] + (i == j ? w->st.sigma : 0.0);
  for (int r = 0; r < m; r++) {
    double rho = w->rho_vec[r];
    for (int a = w->a_rp[r]; a < w->a_rp[r + 1]; a++) {
      double va = rho * w->a_v[a]; int i = w->a_ci[a];
      for (int b = w->a_rp[r]; b < w->a_rp[r + 1]; b++) S[(size_t)i * n + w->a_ci[b]] += va * w->a_v[b];
    }
  }
  /* in-place LDL' (row-oriented, unit lower) */
  for (int j = 0; j < n; j++) {
    double d = S[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= S[(size_t)j * n + k] * S[(size_t)j * n + k] * w->Dl[k];
    if (!(d > 0.0)) return 1;
    w->Dl[j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = S[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= S[(size_t)i * n + k] * S[(size_t)j * n + k] * w->Dl[k];
      S[(size_t)i * n + j] = s / d;
    }
  }
  return 0;
}

/* solve S v = b in place */
static void ldl_solve(const orc_solver *w, double *b) {
  int n = w->n; const double *L = w->L;
  for (int i = 0; i < n; i++) { double s = b[i]; const double *r = L + (size_t)i * n; for (int k = 0; k < i; k++) s -= r[k] * b[k]; b[i] = s; }
  for (int i = 0; i < n; i++) b[i] /= w->Dl[i];
  for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * b[k]; b[i] = s; }
}

orc_solver *orc_setup(int n, int m, const double *P, const double *q, const double *A,
                      const double *l, const double *u, const orc_settings *s) {
  if (n <= 0 || m < 0) return NULL;
  for (int i = 0; i < m; i++) { double li = l ? l[i] : -INFINITY, ui = u ? u[i] : INFINITY; if (li > ui) return NULL; }
  orc_solver *w = (orc_solver *)calloc(1, sizeof(orc_solver));
  w->n = n; w->m = m; w->st = *s;
  w->P = dalloc((size_t)n * n); w->A = dalloc((size_t)m * n);
  w->q = dalloc(n); w->l = dalloc(m); w->u = dalloc(m);
  w->D = dalloc(n); w->Dinv = dalloc(n); w->E = dalloc(m); w->Einv = dalloc(m);
  w->rho_vec = dalloc(m); w->rho_inv_vec = dalloc(m); w->constr_type = (int *)calloc(m ? m : 1, sizeof(int));
  w->rho_vec0 = dalloc(m); w->rho_inv_vec0 = dalloc(m); w->constr_type0 = (int *)calloc(m ? m : 1, sizeof(int));
  w->L = dalloc((size_t)n * n); w->Dl = dalloc(n); w->L0 = dalloc((size_t)n * n); w->Dl0 = dalloc(n);
  w->x = dalloc(n); w->z = dalloc(m); w->y = dalloc(m); w->x_prev = dalloc(n); w->z_prev = dalloc(m);
  w->xt = dalloc(n); w->zt = dalloc(m); w->delta_x = dalloc(n); w->delta_y = dalloc(m);
  w->Ax = dalloc(m); w->Px = dalloc(n); w->Aty = dalloc(n); w->Adelta_x = dalloc(m); w->Atdelta_y = dalloc(n); w->Pdelta_x = dalloc(n);
  w->tn = dalloc(n); w->tm = dalloc(m); w->sol_x = dalloc(n); w->sol_y = dalloc(m);
  /* P: upper triangle mirrored (osqp-eigen hands OSQP triangularView<Upper>) */
  for (int i = 0; i < n; i++) for (int j = i; j < n; j++) { double v = P[(size_t)i * n + j]; w->P[(size_t)i * n + j] = v; w->P[(size_t)j * n + i] = v; }
  memcpy(w->A, A, sizeof(double) * (size_t)m * n);
  for (int j = 0; j < n; j++) w->q[j] = q ? q[j] : 0.0;
  for (int i = 0; i < m; i++) { w->l[i] = l ? l[i] : -INFINITY; w->u[i] = u ? u[i] : INFINITY; }
  if (w->st.scaling) scale_data(w);
  else { w->c = w->cinv = 1.0; for (int j = 0; j < n; j++) w->D[j] = w->Dinv[j] = 1.0; for (int i = 0; i < m; i++) w->E[i] = w->Einv[i] = 1.0; }
  build_csr(w);
  set_rho_vec(w);
  w->rho0 = w->st.rho;
  if (factorize(w)) { orc_cleanup(w); return NULL; }
  memcpy(w->L0, w->L, sizeof(double) * (size_t)n * n); memcpy(w->Dl0, w->Dl, sizeof(double) * n);
  memcpy(w->rho_vec0, w->rho_vec, sizeof(double) * m); memcpy(w->rho_inv_vec0, w->rho_inv_vec, sizeof(double) * m);
  memcpy(w->constr_type0, w->constr_type, sizeof(int) * m);
  w->status_val = ORC_UNSOLVED; w->iter = 0; w->rho_updates = 0;
  return w;
}

void orc_cleanup(orc_solver *w) {
  if (!w) return;
  free(w->P); free(w->A); free(w->q); free(w->l); free(w->u);
  free(w->a_rp); free(w->a_ci); free(w->a_v); free(w->at_rp); free(w->at_ci); free(w->at_v);
  free(w->D); free(w->Dinv); free(w->E); free(w->Einv);
  free(w->rho_vec); free(w->rho_inv_vec); free(w->constr_type);
  free(w->rho_vec0); free(w->rho_inv_vec0); free(w->constr_type0);
  free(w->L); free(w->Dl); free(w->L0); free(w->Dl0);
  free(w->x); free(w->z); free(w->y); free(w->x_prev); free(w->z_prev); free(w->xt); free(w->zt);
  free(w->delta_x); free(w->delta_y); free(w->Ax); free(w->Px); free(w->Aty);
  free(w->Adelta_x); free(w->Atdelta_y); free(w->Pdelta_x); free(w->tn); free(w->tm);
  free(w->sol_x); free(w->sol_y); free(w);
}

static void reset_info(orc_solver *w) { w->status_val = ORC_UNSOLVED; w->rho_updates = 0; }

/* osqp/src/osqp.c: osqp_update_lin_cost */
int orc_update_lin_cost(orc_solver *w, const double *q) {
  for (int j = 0; j < w->n; j++) w->q[j] = q[j];
  if (w->st.scaling) for (int j = 0; j < w->n; j++) w->q[j] = w->c * (w->D[j] * w->q[j]);
  reset_info(w); return 0;
}

/* osqp/src/auxil.c: update_rho_vec (re-classify rows after a bound update) */
static int update_rho_vec(orc_solver *w) {
  int changed = 0;
  for (int i = 0; i < w->m; i++) {
    int t; double r;
    if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) { t = -1; r = RHO_MIN; }
    else if (w->u[i] - w->l[i] < RHO_TOL) { t = 1; r = RHO_EQ_OVER_RHO_INEQ * w->st.rho; }
    else { t = 0; r = w->st.rho; }
    if (t != w->constr_type[i]) { w->constr_type[i] = t; w->rho_vec[i] = r; w->rho_inv_vec[i] = 1.0 / r; changed = 1; }
  }
  if (changed) return factorize(w);
  return 0;
}

int orc_update_bounds(orc_solver *w, const double *l, const double *u) {
  for (int i = 0; i < w->m; i++) if (l[i] > u[i]) return 1;
  for (int i = 0; i < w->m; i++) { w->l[i] = l[i]; w->u[i] = u[i]; }
  if (w->st.scaling) for (int i = 0; i < w->m; i++) { w->l[i] = w->E[i] * w->l[i]; w->u[i] = w->E[i] * w->u[i]; }
  reset_info(w); return update_rho_vec(w);
}
int orc_update_lower_bound(orc_solver *w, const double *l) {
  for (int i = 0; i < w->m; i++) w->l[i] = l[i];
  if (w->st.scaling) for (int i = 0; i < w->m; i++) w->l[i] = w->E[i] * w->l[i];
  for (int i = 0; i < w->m; i++) if (w->l[i] > w->u[i]) return 1;
  reset_info(w); return update_rho_vec(w);
}
int orc_update_upper_bound(orc_solver *w, const double *u) {
  for (int i = 0; i < w->m; i++) w->u[i] = u[i];
  if (w->st.scaling) for (int i = 0; i < w->m; i++) w->u[i] = w->E[i] * w->u[i];
  for (int i = 0; i < w->m; i++) if (w->l[i] > w->u[i]) return 1;
  reset_info(w); return update_rho_vec(w);
}

void orc_cold_start(orc_solver *w) {
  memset(w->x, 0, sizeof(double) * w->n); memset(w->z, 0, sizeof(double) * w->m); memset(w->y, 0, sizeof(double) * w->m);
}

/* osqp/src/osqp.c: osqp_warm_start (x, y given unscaled; z = A x) */
int orc_warm_start(orc_solver *w, const double *x, const double *y) {
  for (int j = 0; j < w->n; j++) w->x[j] = x[j];
  for (int i = 0; i < w->m; i++) w->y[i] = y[i];
  if (w->st.scaling) {
    for (int j = 0; j < w->n; j++) w->x[j] = w->Dinv[j] * w->x[j];
    for (int i = 0; i < w->m; i++) w->y[i] = w->c * (w->Einv[i] * w->y[i]);
  }
  csr_mv(w->m, w->a_rp, w->a_ci, w->a_v, w->x, w->z);
  return 0;
}

void orc_reset(orc_solver *w) {
  orc_cold_start(w);
  w->st.rho = w->rho0;
  memcpy(w->L, w->L0, sizeof(double) * (size_t)w->n * w->n); memcpy(w->Dl, w->Dl0, sizeof(double) * w->n);
  memcpy(w->rho_vec, w->rho_vec0, sizeof(double) * w->m); memcpy(w->rho_inv_vec, w->rho_inv_vec0, sizeof(double) * w->m);
  memcpy(w->constr_type, w->constr_type0, sizeof(int) * w->m);
  reset_info(w);
}

/* osqp/src/auxil.c: compute_pri_res / compute_dua_res / tolerances (update_info) */
static void update_info(orc_solver *w, int iter) {
  int n = w->n, m = w->m;
  w->iter = iter;
  /* objective: 0.5 x'Px + q'x  (scaled), times cinv */
  dense_sym_mv(n, w->P, w->x, w->Px);
  double obj = 0; for (int j = 0; j < n; j++) obj += 0.5 * w->x[j] * w->Px[j] + w->q[j] * w->x[j];
  w->obj_val = w->st.scaling ? w->cinv * obj : obj;
  /* primal residual: z_prev <- A x - z */
  csr_mv(m, w->a_rp, w->a_ci, w->a_v, w->x, w->Ax);
  for (int i = 0; i < m; i++) w->z_prev[i] = w->Ax[i] - w->z[i];
  if (m == 0) w->pri_res = 0;
  else if (w->st.scaling && !w->st.scaled_termination) w->pri_res = scaled_norm_inf(w->Einv, w->z_prev, m);
  else w->pri_res = norm_inf(w->z_prev, m);
  /* dual residual: x_prev <- q + P x + A' y */
  csr_mv(n, w->at_rp, w->at_ci, w->at_v, w->y, w->Aty);
  for (int j = 0; j < n; j++) w->x_prev[j] = (w->q[j] + w->Px[j]) + w->Aty[j];
  if (w->st.scaling && !w->st.scaled_termination) w->dua_res = w->cinv * scaled_norm_inf(w->Dinv, w->x_prev, n);
  else w->dua_res = norm_inf(w->x_prev, n);
}

static double compute_pri_tol(const orc_solver *w, double eps_abs, double eps_rel) {
  double mx;
  if (w->st.scaling && !w->st.scaled_termination) mx = c_max(scaled_norm_inf(w->Einv, w->z, w->m), scaled_norm_inf(w->Einv, w->Ax, w->m));
  else mx = c_max(norm_inf(w->z, w->m), norm_inf(w->Ax, w->m));
  return eps_abs + eps_rel * mx;
}
static double compute_dua_tol(const orc_solver *w, double eps_abs, double eps_rel) {
  double mx;
  if (w->st.scaling && !w->st.scaled_termination) {
    mx = c_max(c_max(scaled_norm_inf(w->Dinv, w->q, w->n), scaled_norm_inf(w->Dinv, w->Aty, w->n)), scaled_norm_inf(w->Dinv, w->Px, w->n));
    mx *= w->cinv;
  } else mx = c_max(c_max(norm_inf(w->q, w->n), norm_inf(w->Aty, w->n)), norm_inf(w->Px, w->n));
  return eps_abs + eps_rel * mx;
}

/* osqp/src/auxil.c: is_primal_infeasible */
static int is_primal_infeasible(orc_solver *w, double eps) {
  int n = w->n, m = w->m; double nd, lhs = 0;
  for (int i = 0; i < m; i++) {
    if (w->u[i] > OSQP_INFTY * MIN_SCALING) {
      if (w->l[i] < -OSQP_INFTY * MIN_SCALING) w->delta_y[i] = 0.0;
      else w->delta_y[i] = c_min(w->delta_y[i], 0.0);
    } else if (w->l[i] < -OSQP_INFTY * MIN_SCALING) w->delta_y[i] = c_max(w->delta_y[i], 0.0);
  }
  if (w->st.scaling && !w->st.scaled_termination) nd = scaled_norm_inf(w->E, w->delta_y, m);
  else nd = norm_inf(w->delta_y, m);
  if (nd > eps) {
    for (int i = 0; i < m; i++) {
      /* after the projection above the infinite side contributes
       * bound*0; written with explicit guards so that -inf*0 never forms a NaN */
      double dp = c_max(w->delta_y[i], 0.0), dm = c_min(w->delta_y[i], 0.0);
      if (dp != 0.0) lhs += w->u[i] * dp;
      if (dm != 0.0) lhs += w->l[i] * dm;
    }
    if (lhs < -eps * nd) {
      csr_mv(n, w->at_rp, w->at_ci, w->at_v, w->delta_y, w->Atdelta_y);
      if (w->st.scaling && !w->st.scaled_termination) for (int j = 0; j < n; j++) w->Atdelta_y[j] *= w->Dinv[j];
      return norm_inf(w->Atdelta_y, n) < eps * nd;
    }
  }
  return 0;
}

/* osqp/src/auxil.c: is_dual_infeasible */
static int is_dual_infeasible(orc_solver *w, double eps) {
  int n = w->n, m = w->m; double nd, cs;
  if (w->st.scaling && !w->st.scaled_termination) { nd = scaled_norm_inf(w->D, w->delta_x, n); cs = w->c; }
  else { nd = norm_inf(w->delta_x, n); cs = 1.0; }
  if (nd > eps) {
    double qd = 0; for (int j = 0; j < n; j++) qd += w->q[j] * w->delta_x[j];
    if (qd < -cs * eps * nd) {
      dense_sym_mv(n, w->P, w->delta_x, w->Pdelta_x);
      if (w->st.scaling && !w->st.scaled_termination) for (int j = 0; j < n; j++) w->Pdelta_x[j] *= w->Dinv[j];
      if (norm_inf(w->Pdelta_x, n) < cs * eps * nd) {
        csr_mv(m, w->a_rp, w->a_ci, w->a_v, w->delta_x, w->Adelta_x);
        if (w->st.scaling && !w->st.scaled_termination) for (int i = 0; i < m; i++) w->Adelta_x[i] *= w->Einv[i];
        for (int i = 0; i < m; i++) {
          if (((w->u[i] < OSQP_INFTY * MIN_SCALING) && (w->Adelta_x[i] > eps * nd)) ||
              ((w->l[i] > -OSQP_INFTY * MIN_SCALING) && (w->Adelta_x[i] < -eps * nd))) return 0;
        }
        return 1;
      }
    }
  }
  return 0;
}

/* osqp/src/auxil.c: check_termination */
static int check_termination(orc_solver *w, int approximate) {
  double eps_abs = w->st.eps_abs, eps_rel = w->st.eps_rel, epi = w->st.eps_prim_inf, edi = w->st.eps_dual_inf;
  int prim_ok = 0, dual_ok = 0, prim_inf = 0, dual_inf = 0;
  if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
  if (w->m == 0) prim_ok = 1;
  else {
    double ep = compute_pri_tol(w, eps_abs, eps_rel);
    if (w->pri_res < ep) prim_ok = 1; else prim_inf = is_primal_infeasible(w, epi);
  }
  double ed = compute_dua_tol(w, eps_abs, eps_rel);
  if (w->dua_res < ed) dual_ok = 1; else dual_inf = is_dual_infeasible(w, edi);
  if (prim_ok && dual_ok) { w->status_val = approximate ? ORC_SOLVED_INACCURATE : ORC_SOLVED; return 1; }
  if (prim_inf) {
    w->status_val = approximate ? ORC_PRIMAL_INFEASIBLE_INACCURATE : ORC_PRIMAL_INFEASIBLE;
    if (w->st.scaling && !w->st.scaled_termination) for (int i = 0; i < w->m; i++) w->delta_y[i] *= w->E[i];
    w->obj_val = OSQP_INFTY; return 1;
  }
  if (dual_inf) {
    w->status_val = approximate ? ORC_DUAL_INFEASIBLE_INACCURATE : ORC_DUAL_INFEASIBLE;
    if (w->st.scaling && !w->st.scaled_termination) for (int j = 0; j < w->n; j++) w->delta_x[j] *= w->D[j];
    w->obj_val = -OSQP_INFTY; return 1;
  }
  return 0;
}

/* osqp/src/auxil.c: compute_rho_estimate (uses the SCALED residual vectors left
 * in z_prev / x_prev by update_info) */
static double compute_rho_estimate(const orc_solver *w) {
  int n = w->n, m = w->m;
  double pri = norm_inf(w->z_prev, m), dua = norm_inf(w->x_prev, n);
  double pn = c_max(norm_inf(w->z, m), norm_inf(w->Ax, m));
  pri /= (pn + OSQP_DIVISION_TOL);
  double dn = c_max(c_max(norm_inf(w->q, n), norm_inf(w->Aty, n)), norm_inf(w->Px, n));
  dua /= (dn + OSQP_DIVISION_TOL);
  double r = w->st.rho * sqrt(pri / (dua + OSQP_DIVISION_TOL));
  return c_min(c_max(r, RHO_MIN), RHO_MAX);
}

/* osqp/src/osqp.c: osqp_update_rho */
static int update_rho(orc_solver *w, double rho_new) {
  w->st.rho = c_min(c_max(rho_new, RHO_MIN), RHO_MAX);
  for (int i = 0; i < w->m; i++) {
    if (w->constr_type[i] == 0) { w->rho_vec[i] = w->st.rho; w->rho_inv_vec[i] = 1.0 / w->st.rho; }
    else if (w->constr_type[i] == 1) { w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->st.rho; w->rho_inv_vec[i] = 1.0 / w->rho_vec[i]; }
  }
  return factorize(w);
}

/* osqp/src/auxil.c: adapt_rho */
static int adapt_rho(orc_solver *w) {
  double rn = compute_rho_estimate(w);
  w->rho_estimate = rn;
  if (rn > w->st.rho * w->st.adaptive_rho_tolerance || rn < w->st.rho / w->st.adaptive_rho_tolerance) {
    int e = update_rho(w, rn); w->rho_updates += 1; return e;
  }
  return 0;
}

/* osqp/src/polish.c: polish() -- refine a SOLVED iterate by solving the KKT system of the active set it identifies.
 *   form_Ared:  row i is lower-active if z_i - l_i < -y_i, upper-active if u_i - z_i < y_i (scaled quantities);
 *               Ared = [lower-active rows; upper-active rows]
 *   solve  [P + delta I, Ared'; Ared, -delta I] [x; y_red] = [-q; l_low; u_upp]  (dense LDL', quasi-definite, no pivoting)
 *   iterative_refinement: polish_refine_iter steps against the UNregularised KKT matrix
 *   z = A x, y expanded to m rows, (z, y) projected onto the normal cone (project_normalcone); update_info on the polished point; accepted (status_polish = 1) only if it
 *   improves the residuals as polish.c tests, otherwise the ADMM solution is kept (status_polish = -1). */
static void polish(orc_solver *w) {
  int n = w->n, m = w->m;
  int *act = (int *)calloc(m ? m : 1, sizeof(int));   /* Ared row -> row of A */
  int *sgn = (int *)calloc(m ? m : 1, sizeof(int));   /* -1 lower-active, +1 upper-active */
  int k = 0;
  for (int i = 0; i < m; i++) if (w->z[i] - w->l[i] < -w->y[i]) { act[k] = i; sgn[k] = -1; k++; }
  for (int i = 0; i < m; i++) if (w->u[i] - w->z[i] < w->y[i]) { act[k] = i; sgn[k] = 1; k++; }
  int N = n + k;
  double *K = dalloc((size_t)N * N), *Dk = dalloc(N), *b = dalloc(N), *sol = dalloc(N), *r = dalloc(N);
  double *px = dalloc(n), *pz = dalloc(m), *py = dalloc(m);
  const double delta = w->st.delta;
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) K[(size_t)i * N + j] = w->P[(size_t)i * n + j] + (i == j ? delta : 0.0);
  for (int a = 0; a < k; a++) {
    for (int j = 0; j < n; j++) { double v = w->A[(size_t)act[a] * n + j]; K[(size_t)(n + a) * N + j] = v; K[(size_t)j * N + n + a] = v; }
    for (int c = 0; c < k; c++) K[(size_t)(n + a) * N + n + c] = a == c ? -delta : 0.0;
  }
  for (int j = 0; j < n; j++) b[j] = -w->q[j];
  for (int a = 0; a < k; a++) b[n + a] = sgn[a] < 0 ? w->l[act[a]] : w->u[act[a]];
  /* in-place LDL' (unit lower L below the diagonal of K, D in Dk) */
  int ok = 1;
  for (int j = 0; j < N && ok; j++) {
    double d = K[(size_t)j * N + j];
    for (int c = 0; c < j; c++) d -= K[(size_t)j * N + c] * K[(size_t)j * N + c] * Dk[c];
    if (d == 0.0 || d != d) { ok = 0; break; }
    Dk[j] = d;
    for (int i = j + 1; i < N; i++) {
      double s = K[(size_t)i * N + j];
      for (int c = 0; c < j; c++) s -= K[(size_t)i * N + c] * K[(size_t)j * N + c] * Dk[c];
      K[(size_t)i * N + j] = s / d;
    }
  }
  w->status_polish = -1;
  if (ok) {
#define KKT_SOLVE(v) do { \
    for (int i = 0; i < N; i++) { double s = (v)[i]; for (int c = 0; c < i; c++) s -= K[(size_t)i * N + c] * (v)[c]; (v)[i] = s; } \
    for (int i = 0; i < N; i++) (v)[i] /= Dk[i]; \
    for (int i = N - 1; i >= 0; i--) { double s = (v)[i]; for (int c = i + 1; c < N; c++) s -= K[(size_t)c * N + i] * (v)[c]; (v)[i] = s; } } while (0)
    memcpy(sol, b, sizeof(double) * N);
    KKT_SOLVE(sol);
    for (int it = 0; it < w->st.polish_refine_iter; it++) {
      /* r = b - [P, Ared'; Ared, 0] sol */
      for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int j = 0; j < n; j++) s -= w->P[(size_t)i * n + j] * sol[j];
        for (int a = 0; a < k; a++) s -= w->A[(size_t)act[a] * n + i] * sol[n + a];
        r[i] = s;
      }
      for (int a = 0; a < k; a++) { double s = b[n + a]; for (int j = 0; j < n; j++) s -= w->A[(size_t)act[a] * n + j] * sol[j]; r[n + a] = s; }
      KKT_SOLVE(r);
      for (int i = 0; i < N; i++) sol[i] += r[i];
    }
#undef KKT_SOLVE
    for (int j = 0; j < n; j++) px[j] = sol[j];
    for (int i = 0; i < m; i++) py[i] = 0.0;
    for (int a = 0; a < k; a++) py[act[a]] = sol[n + a];
    csr_mv(m, w->a_rp, w->a_ci, w->a_v, px, pz);
    /* project_normalcone (proj.c): z <- proj(z + y), y <- (z + y) - proj(z + y); then update_info on the polished point:
     * the primal residual is A x - z */
    double *Ax = dalloc(m), *Px = dalloc(n), *Aty = dalloc(n), *rp = dalloc(m), *rd = dalloc(n);
    for (int i = 0; i < m; i++) {
      Ax[i] = pz[i];
      double zy = pz[i] + py[i];
      pz[i] = c_min(c_max(zy, w->l[i]), w->u[i]);
      py[i] = zy - pz[i];
      rp[i] = Ax[i] - pz[i];
    }
    dense_sym_mv(n, w->P, px, Px);
    csr_mv(n, w->at_rp, w->at_ci, w->at_v, py, Aty);
    double obj = 0; for (int j = 0; j < n; j++) { obj += 0.5 * px[j] * Px[j] + w->q[j] * px[j]; rd[j] = (w->q[j] + Px[j]) + Aty[j]; }
    int unscale = w->st.scaling && !w->st.scaled_termination;
    double pri = m == 0 ? 0.0 : (unscale ? scaled_norm_inf(w->Einv, rp, m) : norm_inf(rp, m));
    double dua = unscale ? w->cinv * scaled_norm_inf(w->Dinv, rd, n) : norm_inf(rd, n);
    int success = (pri < w->pri_res && dua < w->dua_res) || (pri < w->pri_res && w->dua_res < 1e-10) || (dua < w->dua_res && w->pri_res < 1e-10);
    if (success) {
      memcpy(w->x, px, sizeof(double) * n); memcpy(w->z, pz, sizeof(double) * m); memcpy(w->y, py, sizeof(double) * m);
      w->obj_val = w->st.scaling ? w->cinv * obj : obj; w->pri_res = pri; w->dua_res = dua;
      w->status_polish = 1;
    }
    free(Ax); free(Px); free(Aty); free(rp); free(rd);
  }
  free(act); free(sgn); free(K); free(Dk); free(b); free(sol); free(r); free(px); free(pz); free(py);
}

/* osqp/src/osqp.c: osqp_solve */
int orc_solve(orc_solver *w) {
  int n = w->n, m = w->m, iter, can_check = 0;
  const double alpha = w->st.alpha, sigma = w->st.sigma;
  if (!w->st.warm_start) orc_cold_start(w);
  w->status_val = ORC_UNSOLVED;
  for (iter = 1; iter <= w->st.max_iter; iter++) {
    { double *t = w->x; w->x = w->x_prev; w->x_prev = t; t = w->z; w->z = w->z_prev; w->z_prev = t; }
    /* update_xz_tilde: rhs = [sigma x_prev - q ; z_prev - y./rho_vec], KKT solve */
    for (int j = 0; j < n; j++) w->xt[j] = sigma * w->x_prev[j] - w->q[j];
    for (int i = 0; i < m; i++) w->zt[i] = w->z_prev[i] - w->rho_inv_vec[i] * w->y[i];
    /* forward elimination of the constraint block: xt += A' (rho_vec .* zt) */
    for (int i = 0; i < m; i++) w->tm[i] = w->rho_vec[i] * w->zt[i];
    csr_mv(n, w->at_rp, w->at_ci, w->at_v, w->tm, w->tn);
    for (int j = 0; j < n; j++) w->xt[j] += w->tn[j];
    ldl_solve(w, w->xt);
    /* back substitution: nu = rho_vec .* (A xt - rhs_z); z_tilde = rhs_z + nu ./ rho_vec */
    csr_mv(m, w->a_rp, w->a_ci, w->a_v, w->xt, w->tm);
    for (int i = 0; i < m; i++) { double nu = w->rho_vec[i] * (w->tm[i] - w->zt[i]); w->zt[i] = w->zt[i] + w->rho_inv_vec[i] * nu; }
    /* update_x */
    for (int j = 0; j < n; j++) { w->x[j] = alpha * w->xt[j] + (1.0 - alpha) * w->x_prev[j]; w->delta_x[j] = w->x[j] - w->x_prev[j]; }
    /* update_z */
    for (int i = 0; i < m; i++) {
      double v = alpha * w->zt[i] + (1.0 - alpha) * w->z_prev[i] + w->rho_inv_vec[i] * w->y[i];
      w->z[i] = c_min(c_max(v, w->l[i]), w->u[i]);
    }
    /* update_y */
    for (int i = 0; i < m; i++) {
      double dy = alpha * w->zt[i] + (1.0 - alpha) * w->z_prev[i] - w->z[i];
      dy *= w->rho_vec[i]; w->delta_y[i] = dy; w->y[i] += dy;
    }
    can_check = w->st.check_termination && (iter % w->st.check_termination == 0);
    if (can_check) { update_info(w, iter); if (check_termination(w, 0)) break; }
    if (w->st.adaptive_rho && w->st.adaptive_rho_interval && (iter % w->st.adaptive_rho_interval == 0)) {
      if (!can_check) update_info(w, iter);
      if (adapt_rho(w)) return 1;
    }
  }
  if (iter > w->st.max_iter) iter = w->st.max_iter; /* loop ran out */
  if (!can_check) { update_info(w, iter); check_termination(w, 0); }
  if (w->status_val == ORC_UNSOLVED) { if (!check_termination(w, 1)) w->status_val = ORC_MAX_ITER_REACHED; }
  w->rho_estimate = compute_rho_estimate(w);
  w->status_polish = 0;
  if (w->st.polish && w->status_val == ORC_SOLVED) polish(w);
  /* store_solution */
  int has_sol = !(w->status_val == ORC_PRIMAL_INFEASIBLE || w->status_val == ORC_PRIMAL_INFEASIBLE_INACCURATE ||
                  w->status_val == ORC_DUAL_INFEASIBLE || w->status_val == ORC_DUAL_INFEASIBLE_INACCURATE);
  if (has_sol) {
    for (int j = 0; j < n; j++) w->sol_x[j] = w->st.scaling ? w->D[j] * w->x[j] : w->x[j];
    for (int i = 0; i < m; i++) w->sol_y[i] = w->st.scaling ? w->cinv * (w->E[i] * w->y[i]) : w->y[i];
  } else {
    for (int j = 0; j < n; j++) w->sol_x[j] = NAN;
    for (int i = 0; i < m; i++) w->sol_y[i] = NAN;
    orc_cold_start(w); /* OSQP cold-starts the iterates after an infeasible solve */
  }
  return 0;
}

void orc_get_solution(const orc_solver *w, double *x, double *y) {
  if (x) memcpy(x, w->sol_x, sizeof(double) * w->n);
  if (y) memcpy(y, w->sol_y, sizeof(double) * w->m);
}
void orc_get_info(const orc_solver *w, double *o) {
  o[0] = w->status_val; o[1] = w->iter; o[2] = w->rho_updates; o[3] = w->st.rho;
  o[4] = w->obj_val; o[5] = w->pri_res; o[6] = w->dua_res; o[7] = w->rho_estimate;
}
int orc_get_status_polish(const orc_solver *w) { return w->status_polish; }
void orc_get_scaling(const orc_solver *w, double *D, double *E, double *c) {
  if (D) memcpy(D, w->D, sizeof(double) * w->n);
  if (E) memcpy(E, w->E, sizeof(double) * w->m);
  if (c) *c = w->c;
}
void orc_get_scaled_data(const orc_solver *w, double *P, double *A) {
  if (P) memcpy(P, w->P, sizeof(double) * (size_t)w->n * w->n);
  if (A) memcpy(A, w->A, sizeof(double) * (size_t)w->m * w->n);
}
void orc_get_iterates(const orc_solver *w, double *x, double *z, double *y) {
  if (x) memcpy(x, w->x, sizeof(double) * w->n);
  if (z) memcpy(z, w->z, sizeof(double) * w->m);
  if (y) memcpy(y, w->y, sizeof(double) * w->m);
}

/* ------------------------------------------------------------------------- */
typedef struct {
  int n, m, B, tid, nthreads, warm;
  const double *P, *A, *l0, *u0, *q, *l, *u; const orc_settings *s;
  double *x, *y; int *status, *iters; double secs;
} batch_arg;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void *batch_worker(void *p) {
  batch_arg *a = (batch_arg *)p;
  int n = a->n, m = a->m;
  orc_solver *w = orc_setup(n, m, a->P, NULL, a->A, a->l0, a->u0, a->s);
  if (!w) { a->secs = -1; return NULL; }
  int lo = (int)((long long)a->B * a->tid / a->nthreads), hi = (int)((long long)a->B * (a->tid + 1) / a->nthreads);
  double info[8];
  double t0 = now_s();
  for (int b = lo; b < hi; b++) {
    if (!a->warm) orc_reset(w);
    orc_update_lin_cost(w, a->q + (size_t)b * n);
    if (a->l && a->u) orc_update_bounds(w, a->l + (size_t)b * m, a->u + (size_t)b * m);
    else if (a->u) orc_update_upper_bound(w, a->u + (size_t)b * m);
    orc_solve(w);
    orc_get_info(w, info);
    if (a->x) orc_get_solution(w, a->x + (size_t)b * n, NULL);
    if (a->y) orc_get_solution(w, NULL, a->y + (size_t)b * m);
    if (a->status) a->status[b] = (int)info[0];
    if (a->iters) a->iters[b] = (int)info[1];
  }
  a->secs = now_s() - t0;
  orc_cleanup(w);
  return NULL;
}

double orc_solve_batch(int n, int m, const double *P, const double *A, const double *l0, const double *u0,
                       const orc_settings *s, int B, const double *q, const double *l, const double *u,
                       int nthreads, double *x, double *y, int *status, int *iters) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > B) nthreads = B > 0 ? B : 1;
  pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
  batch_arg *args = (batch_arg *)calloc(nthreads, sizeof(batch_arg));
  for (int t = 0; t < nthreads; t++) {
    batch_arg *a = &args[t];
    a->n = n; a->m = m; a->B = B; a->tid = t; a->nthreads = nthreads; a->warm = 0;
    a->P = P; a->A = A; a->l0 = l0; a->u0 = u0; a->q = q; a->l = l; a->u = u; a->s = s;
    a->x = x; a->y = y; a->status = status; a->iters = iters;
    pthread_create(&th[t], NULL, batch_worker, a);
  }
  double secs = 0;
  for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); if (args[t].secs < 0) secs = -1; else if (secs >= 0 && args[t].secs > secs) secs = args[t].secs; }
  free(th); free(args);
  return secs;
}
