/*
 * oracle/osqp_port.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, double) of the OSQP 0.6.x ADMM algorithm that the
 * reference reaches through osqp-eigen (reference call sites:
 * src/ModelPredictiveControlAPI.cpp:51-64,96,99,102,105; member type
 * include/ModelPredictiveControlAPI.h:144).  OSQP / osqp-eigen / QDLDL are
 * third-party, un-vendored and unpinned in the reference
 * (CMakeLists.txt:10 `find_package(OsqpEigen REQUIRED)`), and not installed in
 * this image, so this file restates the published algorithm (Stellato et al.
 * 2020, OSQP 0.6.x sources) -- see SURVEY.md section 3.4.
 *
 * PARITY STATUS: "parity unpinned" against real osqp-eigen (no golden vectors
 * exist in the reference and the library is absent).  The port is pinned
 * instead by the solver-independent exact KKT known answers of SURVEY.md
 * section 8(c) (tests/test_oracle_known_answers.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may link or call anything in oracle/.
 */
#ifndef ORACLE_OSQP_PORT_H
#define ORACLE_OSQP_PORT_H

#ifdef __cplusplus
extern "C" {
#endif

/* OSQP status values (osqp/include/constants.h, 0.6.x) */
#define ORC_SOLVED 1
#define ORC_SOLVED_INACCURATE 2
#define ORC_PRIMAL_INFEASIBLE_INACCURATE 3
#define ORC_DUAL_INFEASIBLE_INACCURATE 4
#define ORC_MAX_ITER_REACHED (-2)
#define ORC_PRIMAL_INFEASIBLE (-3)
#define ORC_DUAL_INFEASIBLE (-4)
#define ORC_UNSOLVED (-10)

typedef struct {
  double rho;                    /* 0.1   */
  double sigma;                  /* 1e-6  */
  double alpha;                  /* 1.6   */
  double eps_abs;                /* 1e-3 in OSQP; north star runs 1e-5 */
  double eps_rel;                /* 1e-3 in OSQP; north star runs 1e-5 */
  double eps_prim_inf;           /* 1e-4  */
  double eps_dual_inf;           /* 1e-4  */
  double adaptive_rho_tolerance; /* 5     */
  int max_iter;                  /* 4000  */
  int check_termination;         /* 25    */
  int scaling;                   /* 10    */
  int adaptive_rho;              /* 1     */
  int adaptive_rho_interval;     /* OSQP: 0 = timing based (irreproducible);
                                    port: fixed, default 25 */
  int warm_start;                /* 1     */
  int scaled_termination;        /* 0     */
  int polish;                    /* 0     (osqp polish.c; the reference leaves it off, cpp:51-52) */
  int polish_refine_iter;        /* 3     */
  double delta;                  /* 1e-6  regularisation of the polish KKT system */
} orc_settings;

typedef struct orc_solver orc_solver;

void orc_default_settings(orc_settings *s);

/* P: n*n row-major (only the upper triangle is read, as osqp-eigen passes
 * triangularView<Upper>); A: m*n row-major; q,l,u may be NULL (-> 0, -inf, +inf).
 * Returns NULL on invalid data (l > u, non-positive dims). */
orc_solver *orc_setup(int n, int m, const double *P, const double *q,
                      const double *A, const double *l, const double *u,
                      const orc_settings *s);
void orc_cleanup(orc_solver *w);

int orc_update_lin_cost(orc_solver *w, const double *q);
int orc_update_bounds(orc_solver *w, const double *l, const double *u);
int orc_update_lower_bound(orc_solver *w, const double *l);
int orc_update_upper_bound(orc_solver *w, const double *u);
int orc_warm_start(orc_solver *w, const double *x, const double *y);
void orc_cold_start(orc_solver *w);
/* back to the state right after orc_setup: x=z=y=0, rho=rho0, initial factor */
void orc_reset(orc_solver *w);

int orc_solve(orc_solver *w); /* returns 0; status via orc_get_info */

void orc_get_solution(const orc_solver *w, double *x, double *y);
/* info[0]=status_val info[1]=iter info[2]=rho_updates info[3]=rho
 * info[4]=obj_val info[5]=pri_res info[6]=dua_res info[7]=rho_estimate */
void orc_get_info(const orc_solver *w, double *info8);
/* polish outcome of the last solve: 0 = not run, 1 = successful, -1 = unsuccessful (osqp info->status_polish) */
int orc_get_status_polish(const orc_solver *w);
/* scaling vectors and the scaled data, for parity tests of the device setup */
void orc_get_scaling(const orc_solver *w, double *D, double *E, double *c);
void orc_get_scaled_data(const orc_solver *w, double *Pbar, double *Abar);
/* internal scaled iterates (x, z, y) -- for warm-start parity */
void orc_get_iterates(const orc_solver *w, double *x, double *z, double *y);

/* "one solver per core" batch driver (SURVEY 8d CPU side).  Every problem is
 * an independent solver set up with q = 0 and bounds (l0,u0) exactly as the
 * reference constructor does (cpp:22-23,38-43,54-64), then updated with its own
 * q / l / u and solved.  q:[B][n]; l,u:[B][m] (l may be NULL = keep l0).
 * warm==0: orc_reset before each problem (independent cold solves).
 * Outputs x:[B][n], y:[B][m], status:[B], iters:[B] (any may be NULL).
 * Returns wall seconds spent in the update+solve loop (max over threads). */
double orc_solve_batch(int n, int m, const double *P, const double *A,
                       const double *l0, const double *u0,
                       const orc_settings *s, int B, const double *q,
                       const double *l, const double *u, int nthreads,
                       double *x, double *y, int *status, int *iters);

#ifdef __cplusplus
}
#endif
#endif
