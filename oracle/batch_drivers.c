/*
 * oracle/batch_drivers.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see osqp_port.h).
 *
 * "One controller per core" CPU drivers of the workloads bench.py times beside the GPU (BASELINE.md section 3:
 * osqp-eigen on the host cores, one solver instance per core).  Each worker thread owns a contiguous slice of the
 * controllers and does for every one of them exactly what one ModelPredictiveControlAPI object does:
 *   constructor   (reference src/ModelPredictiveControlAPI.cpp:3-65):   orc_mpc_build + orc_setup
 *   controllerStep (cpp:81-108): setF (cpp:374), upper bound (cpp:99), updateGradient / updateUpperBound, solve, U += dU[0]
 * Plain C on top of osqp_port.c / mpc_assembly.c; no algorithm of its own.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mpc_assembly.h"
#include "osqp_port.h"

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

typedef struct {
  double *H, *Gbar, *Fx, *Fu, *Fr, *Sbar, *Ku, *W0, *Sx, *Su, *CAB, *lb, *f, *ub, *ref, *x;
} mpc_mats;

static void mats_alloc(mpc_mats *t, int N, int nx) {
  size_t n = (size_t)N;
  t->H = calloc(n * n, 8); t->Gbar = calloc(2 * n * n, 8); t->Fx = calloc(n * nx, 8); t->Fu = calloc(n, 8);
  t->Fr = calloc(n * n, 8); t->Sbar = calloc(2 * n * nx, 8); t->Ku = calloc(2 * n, 8); t->W0 = calloc(2 * n, 8);
  t->Sx = calloc(n * nx, 8); t->Su = calloc(n * n, 8); t->CAB = calloc(n, 8); t->lb = calloc(2 * n, 8);
  t->f = calloc(n, 8); t->ub = calloc(2 * n, 8); t->ref = calloc(n, 8); t->x = calloc(n, 8);
  for (size_t i = 0; i < 2 * n; i++) t->lb[i] = orc_mpc_lower_bound();
}
static void mats_free(mpc_mats *t) {
  free(t->H); free(t->Gbar); free(t->Fx); free(t->Fu); free(t->Fr); free(t->Sbar); free(t->Ku); free(t->W0);
  free(t->Sx); free(t->Su); free(t->CAB); free(t->lb); free(t->f); free(t->ub); free(t->ref); free(t->x);
}

/* ------------------------------------------------------------------ per-plant controllers (BASELINE config 4) */
typedef struct {
  int N, nx, B, tid, nthreads, n_state_rows;
  const double *Ad, *Bd, *Cd, *K, *X, *U, *ref; double Q, R, RD, u_limit; const orc_settings *s;
  double *x; int *status, *iters; double secs;
} plant_arg;

static void *plant_worker(void *p) {
  plant_arg *a = (plant_arg *)p;
  const int N = a->N, nx = a->nx;
  mpc_mats t; mats_alloc(&t, N, nx);
  int lo = (int)((long long)a->B * a->tid / a->nthreads), hi = (int)((long long)a->B * (a->tid + 1) / a->nthreads);
  double info[8];
  double t0 = now_s();
  for (int b = lo; b < hi; b++) {
    /* constructor: builders + solver setup with q = 0, l = -DBL_MAX, u = W0 (cpp:22-23,38-43,54-64) */
    orc_mpc_build(N, nx, a->Ad + (size_t)b * nx * nx, a->Bd + (size_t)b * nx, a->Cd, a->K, a->Q, a->R, a->RD, a->n_state_rows,
                  a->u_limit, t.H, t.Gbar, t.Fx, t.Fu, t.Fr, t.Sbar, t.Ku, t.W0, t.Sx, t.Su, t.CAB);
    orc_solver *w = orc_setup(N, 2 * N, t.H, NULL, t.Gbar, t.lb, t.W0, a->s);
    if (!w) { if (a->status) a->status[b] = ORC_UNSOLVED; continue; }
    /* controllerStep */
    for (int j = 0; j < N; j++) t.ref[j] = a->ref[b];
    orc_mpc_step_vectors(N, nx, t.Fx, t.Fu, t.Fr, t.Sbar, t.Ku, t.W0, a->X + (size_t)b * nx, a->U[b], t.ref, t.f, t.ub);
    orc_update_lin_cost(w, t.f); orc_update_upper_bound(w, t.ub);
    orc_solve(w);
    orc_get_info(w, info);
    if (a->x) orc_get_solution(w, a->x + (size_t)b * N, NULL);
    if (a->status) a->status[b] = (int)info[0];
    if (a->iters) a->iters[b] = (int)info[1];
    orc_cleanup(w);
  }
  a->secs = now_s() - t0;
  mats_free(&t);
  return NULL;
}

/* Ad:[B][nx][nx], Bd:[B][nx], X:[B][nx], U:[B], ref:[B]; outputs x:[B][N], status:[B], iters:[B] (may be NULL).
 * Returns wall seconds (max over threads) of constructor + controllerStep for every controller. */
double orc_plant_batch(int N, int nx, int B, const double *Ad, const double *Bd, const double *Cd, const double *K, double Q,
                       double R, double RD, int n_state_rows, double u_limit, const double *X, const double *U,
                       const double *ref, const orc_settings *s, int nthreads, double *x, int *status, int *iters) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > B) nthreads = B > 0 ? B : 1;
  pthread_t *th = calloc(nthreads, sizeof(pthread_t));
  plant_arg *args = calloc(nthreads, sizeof(plant_arg));
  for (int t = 0; t < nthreads; t++) {
    plant_arg *a = &args[t];
    a->N = N; a->nx = nx; a->B = B; a->tid = t; a->nthreads = nthreads; a->n_state_rows = n_state_rows;
    a->Ad = Ad; a->Bd = Bd; a->Cd = Cd; a->K = K; a->X = X; a->U = U; a->ref = ref; a->Q = Q; a->R = R; a->RD = RD;
    a->u_limit = u_limit; a->s = s; a->x = x; a->status = status; a->iters = iters;
    pthread_create(&th[t], NULL, plant_worker, a);
  }
  double secs = 0;
  for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); if (args[t].secs > secs) secs = args[t].secs; }
  free(th); free(args);
  return secs;
}

/* ------------------------------------------------------------------ warm-started closed loop (BASELINE config 5 / 1) */
typedef struct {
  int N, nx, B, tid, nthreads, n_state_rows, steps, skip, period;
  const double *Ad, *Bd, *Cd, *K; double Q, R, RD, u_limit, amp; const int *phase; const orc_settings *s;
  double *X, *U; long long not_solved, iterations; double secs; double *step_secs;
} loop_arg;

static void *loop_worker(void *p) {
  loop_arg *a = (loop_arg *)p;
  const int N = a->N, nx = a->nx;
  int lo = (int)((long long)a->B * a->tid / a->nthreads), hi = (int)((long long)a->B * (a->tid + 1) / a->nthreads);
  int cnt = hi - lo;
  mpc_mats t; mats_alloc(&t, N, nx);
  /* shared plant: one set of builder outputs per thread, one solver per controller (warm start and rho persist, cpp:52) */
  orc_mpc_build(N, nx, a->Ad, a->Bd, a->Cd, a->K, a->Q, a->R, a->RD, a->n_state_rows, a->u_limit, t.H, t.Gbar, t.Fx, t.Fu,
                t.Fr, t.Sbar, t.Ku, t.W0, t.Sx, t.Su, t.CAB);
  orc_solver **w = calloc(cnt > 0 ? cnt : 1, sizeof(orc_solver *));
  for (int i = 0; i < cnt; i++) w[i] = orc_setup(N, 2 * N, t.H, NULL, t.Gbar, t.lb, t.W0, a->s);
  double info[8], xn[64];
  a->secs = 0; a->not_solved = 0; a->iterations = 0;
  for (int k = 0; k < a->steps; k++) {
    double t0 = now_s();
    for (int i = 0; i < cnt; i++) {
      int b = lo + i;
      double *X = a->X + (size_t)b * nx;
      double r = a->amp;
      if (a->period) { int ph = (k + (a->phase ? a->phase[b] : 0)) % a->period; r = (2 * ph < a->period) ? a->amp : -a->amp; }
      for (int j = 0; j < N; j++) t.ref[j] = r;
      orc_mpc_step_vectors(N, nx, t.Fx, t.Fu, t.Fr, t.Sbar, t.Ku, t.W0, X, a->U[b], t.ref, t.f, t.ub);
      orc_update_lin_cost(w[i], t.f); orc_update_upper_bound(w[i], t.ub);
      orc_solve(w[i]);
      orc_get_info(w[i], info);
      if ((int)info[0] == ORC_SOLVED) { orc_get_solution(w[i], t.x, NULL); a->U[b] += t.x[0]; }   /* cpp:102-105 */
      else if (k >= a->skip) a->not_solved++;
      if (k >= a->skip) a->iterations += (long long)info[1];
      /* synthetic plant: X <- Ad X + Bd U */
      for (int rr = 0; rr < nx; rr++) { double s = 0; for (int c = 0; c < nx; c++) s += a->Ad[rr * nx + c] * X[c]; xn[rr] = s + a->Bd[rr] * a->U[b]; }
      memcpy(X, xn, sizeof(double) * nx);
    }
    double dt = now_s() - t0;
    if (k >= a->skip) a->secs += dt;
    if (a->step_secs && a->tid == 0) a->step_secs[k] = dt;
  }
  for (int i = 0; i < cnt; i++) orc_cleanup(w[i]);
  free(w);
  mats_free(&t);
  return NULL;
}

/* `steps` closed-loop steps of B controllers sharing the plant (nx <= 64); the first `skip` steps (the cold solve) are run
 * but not timed / counted.  X:[B][nx], U:[B] are updated in place; phase:[B] or NULL; period == 0: constant reference amp.
 * step_secs: [steps] wall time of every step of thread 0 (may be NULL; the latency probe runs B = 1).
 * Returns the wall seconds (max over threads) of the timed steps. */
double orc_closed_loop(int N, int nx, int B, const double *Ad, const double *Bd, const double *Cd, const double *K, double Q,
                       double R, double RD, int n_state_rows, double u_limit, double *X, double *U, double amp, int period,
                       const int *phase, int steps, int skip, const orc_settings *s, int nthreads, long long *not_solved,
                       long long *iterations, double *step_secs) {
  if (nx > 64) return -1.0;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > B) nthreads = B > 0 ? B : 1;
  pthread_t *th = calloc(nthreads, sizeof(pthread_t));
  loop_arg *args = calloc(nthreads, sizeof(loop_arg));
  for (int t = 0; t < nthreads; t++) {
    loop_arg *a = &args[t];
    a->N = N; a->nx = nx; a->B = B; a->tid = t; a->nthreads = nthreads; a->n_state_rows = n_state_rows; a->steps = steps;
    a->skip = skip; a->period = period; a->Ad = Ad; a->Bd = Bd; a->Cd = Cd; a->K = K; a->Q = Q; a->R = R; a->RD = RD;
    a->u_limit = u_limit; a->amp = amp; a->phase = phase; a->s = s; a->X = X; a->U = U; a->step_secs = step_secs;
    pthread_create(&th[t], NULL, loop_worker, a);
  }
  double secs = 0; long long bad = 0, it = 0;
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    if (args[t].secs > secs) secs = args[t].secs;
    bad += args[t].not_solved; it += args[t].iterations;
  }
  if (not_solved) *not_solved = bad;
  if (iterations) *iterations = it;
  free(th); free(args);
  return secs;
}
