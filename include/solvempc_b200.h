/*
 * solvempc_b200.h -- C ABI of the B200-native batched MPC QP solver.
 *
 * This is the drop-in boundary for the hot path of LukeSchmitt96/solveMPC
 * (SURVEY.md section 8b).  Plain C: POD structs, raw pointers, int return codes
 * (0 = ok), so it is immune to the reference's -D_GLIBCXX_USE_CXX11_ABI=0
 * (reference CMakeLists.txt:22).  Every entry point names the reference
 * interface it replaces.  All device work is CUDA for sm_100a; there is no CPU
 * fallback: without a CUDA device every compute call returns SMPC_ERR_CUDA.
 *
 * Two layers, mirroring the reference's own layering:
 *   smpc_solver_*  replaces the OsqpEigen::Solver member
 *                  (reference include/ModelPredictiveControlAPI.h:144) for a
 *                  BATCH of independent QPs  min .5 x'Px + q'x  s.t. l <= Ax <= u
 *                  solved by the OSQP ADMM iteration on the device;
 *   smpc_mpc_*     replaces class ModelPredictiveControlAPI
 *                  (reference include/ModelPredictiveControlAPI.h:47-243,
 *                  src/ModelPredictiveControlAPI.cpp) for a batch of controllers.
 *
 * Array conventions: dense matrices row-major doubles; batched vectors are
 * [batch][len] contiguous.  `loc` says where a caller buffer lives.
 * Threading: a handle is used from one host thread at a time (the reference is
 * single threaded); different handles are independent.
 */
#ifndef SOLVEMPC_B200_H
#define SOLVEMPC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SMPC_HOST 0
#define SMPC_DEVICE 1

/* return codes */
#define SMPC_OK 0
#define SMPC_ERR_ARG 1      /* bad argument / shape (osqp-eigen returns false) */
#define SMPC_ERR_DATA 2     /* invalid QP data: l > u, P+sigma*I+rho*A'A not PD */
#define SMPC_ERR_CUDA 3     /* CUDA runtime error or no device: hard error, no fallback */
#define SMPC_ERR_STATE 4    /* call order violated (e.g. solve before setup) */
#define SMPC_ERR_IO 5       /* config file missing / malformed (reference throws, cpp:13,437-480) */

/* per-problem status, numbered as OSQP numbers them (osqp constants.h) */
#define SMPC_SOLVED 1
#define SMPC_SOLVED_INACCURATE 2
#define SMPC_PRIMAL_INFEASIBLE_INACCURATE 3
#define SMPC_DUAL_INFEASIBLE_INACCURATE 4
#define SMPC_MAX_ITER_REACHED (-2)
#define SMPC_PRIMAL_INFEASIBLE (-3)
#define SMPC_DUAL_INFEASIBLE (-4)
#define SMPC_UNSOLVED (-10)

/* OSQP settings the reference reaches through solver.settings()
 * (src/ModelPredictiveControlAPI.cpp:51-52 sets verbosity and warm start only;
 * everything else is the OSQP default).  adaptive_rho_interval is FIXED here
 * (OSQP's 0 = "timing based" is irreproducible): default 25. */
typedef struct smpc_settings {
  double rho, sigma, alpha;
  double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
  double adaptive_rho_tolerance;
  int max_iter, check_termination, scaling;
  int adaptive_rho, adaptive_rho_interval;
  int warm_start, scaled_termination;
  int kernel;               /* 0 = auto; 1 = generic warp kernel; 2 = register-resident small-QP kernel (n <= 16, m <= 32; auto choice there);
                               4 = DMMA tile kernel (8/16 QPs per CTA, FP64 tensor pipe; auto choice for n > 16);
                               5 = DMMA variant of the small-QP kernel (8 QPs per 4-warp CTA; higher throughput ceiling at very
                                   large batches, longer single-iteration latency than 2) */
} smpc_settings;

void smpc_default_settings(smpc_settings *s);   /* OSQP 0.6 defaults (eps 1e-3) */
const char *smpc_last_error(void);              /* thread-local message for the last non-zero return */
const char *smpc_version(void);
int smpc_device_count(void);                    /* 0 when no CUDA device is usable */

/* ------------------------------------------------------------------ QP layer */
typedef struct smpc_solver smpc_solver;

/* Replaces solver.data()->setNumberOfVariables/Constraints/HessianMatrix/
 * LinearConstraintsMatrix/LowerBound/UpperBound + solver.initSolver()
 * (cpp:54-64) for `batch` solvers that SHARE P and A ("shared-factor" regime).
 * Every solver is set up with the gradient q0 (length n; NULL = 0 -- the reference
 * constructor passes f(X=U=ref=0) = 0, cpp:22-23,38-39,58; OSQP's cost scaling c depends
 * on it) and the bounds (l0,u0) (length m; NULL = -inf / +inf); the bounds also fix the
 * row classification (equality / inequality / free) used for rho_vec.
 * Only the upper triangle of P is read (osqp-eigen passes triangularView<Upper>). */
int smpc_solver_create_shared(smpc_solver **out, int device, int n, int m, int batch,
                              const double *P, const double *A, const double *q0,
                              const double *l0, const double *u0,
                              const smpc_settings *settings);
/* Same with P (upper triangle) and A in compressed-sparse-column form, the
 * layout Eigen::SparseMatrix<double> / OSQP csc hold (cpp:57,59). */
int smpc_solver_create_shared_csc(smpc_solver **out, int device, int n, int m, int batch,
                                  const int *P_colptr, const int *P_rowidx, const double *P_val,
                                  const int *A_colptr, const int *A_rowidx, const double *A_val,
                                  const double *q0, const double *l0, const double *u0,
                                  const smpc_settings *settings);
/* `batch` solvers with their OWN P_i, A_i ("per-instance" regime, batched Cholesky).
 * P:[batch][n][n], A:[batch][m][n] at `loc`; l0,u0 as above (shared, length m). */
int smpc_solver_create_batched(smpc_solver **out, int device, int n, int m, int batch,
                               const double *P, const double *A, int loc,
                               const double *l0, const double *u0,
                               const smpc_settings *settings);
int smpc_solver_destroy(smpc_solver *s);
int smpc_solver_set_stream(smpc_solver *s, void *cuda_stream);   /* cudaStream_t; NULL = default */
int smpc_solver_dims(const smpc_solver *s, int *n, int *m, int *batch);

/* solver.updateGradient(f) (cpp:96) -> osqp_update_lin_cost, for every instance. q:[batch][n] */
int smpc_solver_update_lin_cost(smpc_solver *s, const double *q, int loc);
/* solver.updateUpperBound(u) (cpp:99) -> osqp_update_upper_bound. u:[batch][m] */
int smpc_solver_update_upper_bound(smpc_solver *s, const double *u, int loc);
int smpc_solver_update_lower_bound(smpc_solver *s, const double *l, int loc);
int smpc_solver_update_bounds(smpc_solver *s, const double *l, const double *u, int loc);
/* osqp_warm_start: x:[batch][n], y:[batch][m] unscaled; z = A x */
int smpc_solver_warm_start(smpc_solver *s, const double *x, const double *y, int loc);
/* x = z = y = 0 (osqp cold_start); rho keeps its adapted value as in OSQP */
int smpc_solver_cold_start(smpc_solver *s);
/* back to the state right after create: x = z = y = 0 and rho = settings.rho */
int smpc_solver_reset(smpc_solver *s);

/* on != 0: every solve starts from the state right after create (x = z = y = 0, rho = settings.rho),
 * i.e. each instance behaves like a freshly constructed solver -- the "independent cold QPs" workload
 * (BASELINE config 2).  Default off: warm start and rho persist across solves as in OSQP (cpp:52). */
int smpc_solver_set_cold_solves(smpc_solver *s, int on);
/* on != 0 (default): the small-QP kernel orders its work queue longest-expected-first with a cheap pre-pass
 * (classify_small_kernel); results do not depend on it */
int smpc_solver_set_scheduling(smpc_solver *s, int on);
/* on != 0: bracket the ADMM kernel of every solve with CUDA events on the handle's stream;
 * smpc_solver_kernel_ms sums and counts them (used by bench.py for the roofline line) */
int smpc_solver_enable_timing(smpc_solver *s, int on);
int smpc_solver_kernel_ms(smpc_solver *s, double *total_ms, int *launches, int reset);

/* solver.settings()->setPolish / setDelta / setPolishRefineIter (osqp-eigen Settings; OSQP polish.c): after the ADMM loop every
 * instance that ended SOLVED guesses its active set from (z, y), solves the regularised reduced KKT system
 * [P + delta I, Ared'; Ared, -delta I] with `refine_iter` steps of iterative refinement and keeps the polished point when it
 * lowers the residuals.  Off by default -- the reference never enables it (cpp:51-52); OSQP defaults delta = 1e-6, 3 steps.
 * One extra kernel launch per solve.  Both regimes. */
int smpc_solver_set_polish(smpc_solver *s, int on, double delta, int refine_iter);
/* info->status_polish of every instance after the last solve: 1 polished, -1 polish unsuccessful (ADMM solution kept),
 * 0 not run (polish off, or the instance did not end SOLVED) */
int smpc_solver_get_polish_status(smpc_solver *s, int *status_polish, int loc);

/* solver.solve() (cpp:102) -> osqp_solve for every instance; asynchronous on the stream */
int smpc_solver_solve(smpc_solver *s);
/* solver.getSolution() (cpp:105): x:[batch][n]; y:[batch][m] duals (either may be NULL).
 * Infeasible instances hold NaN, as OSQP stores. Synchronises when loc == SMPC_HOST. */
int smpc_solver_get_solution(smpc_solver *s, double *x, double *y, int loc);
/* per-instance info (any pointer may be NULL): status (SMPC_* above), iterations,
 * objective, unscaled residuals, final rho, number of rho updates */
int smpc_solver_get_info(smpc_solver *s, int *status, int *iter, double *obj,
                         double *pri_res, double *dua_res, double *rho, int *rho_updates, int loc);
/* number of instances whose status == SMPC_SOLVED after the last solve (host sync);
 * osqp-eigen's solve() returns true exactly for that status */
int smpc_solver_count_solved(smpc_solver *s, int *count);
int smpc_solver_sync(smpc_solver *s);
/* scaling computed at setup (D:n, E:m, c) -- for parity tests of the device-side setup */
int smpc_solver_get_scaling(smpc_solver *s, double *D, double *E, double *c);
/* kernels launched by this handle since creation (bench.py gpu_launches) */
long long smpc_solver_launch_count(const smpc_solver *s);
const char *smpc_solver_kernel_name(const smpc_solver *s);
/* m / 2 when the constraint rows come as pairs [G; -G] (the reference's two-sided limit, cpp:335) AND the handle's
 * kernel exploits it (tile kernel: iteration GEMMs on the top half only, identical iterates); 0 otherwise */
int smpc_solver_row_pairs(const smpc_solver *s);

/* Host-only inspection of the shared-factor plan built at create time (Ruiz scaling D, E, c as
 * OSQP scale_data computes them; the pencil decomposition V, lambda; the operators the kernels
 * use).  Needs no device; any output may be NULL.  Matrices row-major: V,SG,PVT,VinvT n*n; W m*n. */
int smpc_shared_plan_inspect(int n, int m, const double *P, const double *A, const double *q0,
                             const double *l0, const double *u0, const smpc_settings *settings, double *D, double *E,
                             double *c, double *lam, double *V, double *SG, double *W, double *PVT,
                             double *VinvT, signed char *ctype);

/* ----------------------------------------------------------------- MPC layer */
typedef struct smpc_mpc smpc_mpc;

/* The values the reference reads from config/MPC_API.json (cpp:16,19,113-116,138-140)
 * with its compile-time dimensions (include/ModelPredictiveControlAPI.h:26-32) lifted
 * to run time.  N_C = N_O = 1 as in the reference. */
typedef struct smpc_mpc_config {
  int horizon;          /* mpcWindow (h:26), 15 in the reference */
  int nx;               /* N_S (h:30), 4 */
  int n_state_rows;     /* rows of S that carry K: literal 10 at cpp:185 */
  double Q, R, RD;      /* cpp:138-140 */
  double u_limit;       /* literal 255.0 at cpp:368 */
  double xref;          /* cpp:19 */
  const double *Ad;     /* nx*nx row-major (cpp:113); per_instance: [batch][nx][nx] */
  const double *Bd;     /* nx (cpp:114);              per_instance: [batch][nx]     */
  const double *Cd;     /* nx (cpp:115) */
  const double *K;      /* nx (cpp:16)  */
  int per_instance;     /* 0: all controllers share the plant; 1: Ad, Bd differ per instance */
} smpc_mpc_config;

/* ModelPredictiveControlAPI::ModelPredictiveControlAPI (cpp:3-65) for `batch` controllers:
 * runs setTransformations/setLL/setH/setFVars/setLinearConstraints/setUpperBound ON THE DEVICE,
 * then sets up the batched solver (X = U = ref = 0 as cpp:22-23,38-43). */
int smpc_mpc_create(smpc_mpc **out, int device, const smpc_mpc_config *cfg, int batch,
                    const smpc_settings *settings);
/* same, reading the reference's MPC_API.json schema (unchanged keys; additive keys
 * "horizon", "n_state_rows", "u_limit" optional) */
int smpc_mpc_create_from_json(smpc_mpc **out, int device, const char *json_path, int batch,
                              const smpc_settings *settings);
int smpc_mpc_destroy(smpc_mpc *m);
int smpc_mpc_set_stream(smpc_mpc *m, void *cuda_stream);
int smpc_mpc_dims(const smpc_mpc *m, int *horizon, int *nx, int *n, int *mrows, int *batch);
smpc_solver *smpc_mpc_solver(smpc_mpc *m);   /* the `solver` member (h:144) */

/* builder outputs by the reference's member name: "H" "Gbar" "Fx" "Fu" "Fr" "Sbar" "Ku" "W0"
 * "Sx" "Su" "CAB" (shared plant) -- or instance `index` when per_instance; row-major. */
int smpc_mpc_get_matrix(smpc_mpc *m, const char *name, int index, double *out, int capacity);

/* the public members main() writes: X (h:186), U (h:187), and the reference signal
 * (updateRef, cpp:378-381).  X:[batch][nx]; U:[batch]; ref:[batch] (held constant over
 * the horizon as updateRef does) -- any may be NULL to keep the current value. */
int smpc_mpc_set_state(smpc_mpc *m, const double *X, const double *U, const double *ref, int loc);
/* controllerStep (cpp:81-108): f and ub from X,U,ref -> updateGradient -> updateUpperBound ->
 * solve -> U += dU[0].  Asynchronous on the stream. */
int smpc_mpc_controller_step(smpc_mpc *m);
/* One iteration of main()'s loop body up to the solve (src/solver.cpp:45-55: write X, U, ref, then controllerStep) in one
 * call: equivalent to smpc_mpc_set_state(m, X, U, ref, loc) + smpc_mpc_controller_step(m), all three arrays required.
 * Device buffers and pinned host buffers are read where they lie by the step's first kernel (no gather launch in front; the
 * controller's own X, U, ref are updated by that kernel); they must stay unchanged until the stream has run the step.
 * Pageable host buffers, per-instance plants and the larger-QP kernels take the two-call path internally. */
int smpc_mpc_controller_step_from(smpc_mpc *m, const double *X, const double *U, const double *ref, int loc);
/* What main() reads after controllerStep (src/solver.cpp:55-60: the flag and U), delivered by the step itself: every later
 * controllerStep also writes U:[batch] (after U += dU[0]) and status:[batch] (OSQP numbers) to these buffers -- device
 * memory or PINNED host memory (pageable host memory is refused: SMPC_ERR_ARG); either may be NULL, both NULL unbinds.  The
 * one-warp small-QP kernel stores them as each instance ends, the other kernels are followed by one export launch.
 * They are valid once the stream has run the step: smpc_mpc_sync.  (smpc_mpc_closed_loop keeps everything on the device and
 * does not write them.) */
int smpc_mpc_bind_results(smpc_mpc *m, double *U, int *status, int loc);
int smpc_mpc_sync(smpc_mpc *m);
/* synthetic plant for closed-loop runs (the reference's plant is hardware): X <- Ad X + Bd U */
int smpc_mpc_plant_step(smpc_mpc *m);
/* Closed-loop driver: the reference's main loop (src/solver.cpp:43-74: read state -> controllerStep -> write U) for the
 * whole batch without leaving the device.  Each of the `steps` iterations sets the reference (ref_period >= 2: square wave
 * +-ref_amplitude with that period in steps and a per-instance `phase` offset, host int[batch] or NULL = 0;
 * ref_period == 0: keep the current reference), runs controllerStep (warm started, rho persistent, cpp:52) and the
 * synthetic plant step X <- Ad X + Bd U.  use_graph != 0 captures one step into a CUDA graph and replays it.
 * Outputs (may be NULL): solves that did not end SOLVED, and the total number of ADMM iterations.  Synchronises. */
int smpc_mpc_closed_loop(smpc_mpc *m, int steps, double ref_amplitude, int ref_period, const int *phase, int use_graph,
                         long long *not_solved, long long *iterations);
int smpc_mpc_get_state(smpc_mpc *m, double *X, double *U, int loc);
/* what the reference's main loop reads after controllerStep (src/solver.cpp:55-60: the flag and U) in ONE transfer + one
 * synchronisation: U:[batch] and the per-instance OSQP status:[batch] (either may be NULL) */
int smpc_mpc_get_control_status(smpc_mpc *m, double *U, int *status, int loc);
/* vectors handed to the solver in the last controllerStep: f:[batch][n], ub:[batch][2N] */
int smpc_mpc_get_step_vectors(smpc_mpc *m, double *f, double *ub, int loc);
long long smpc_mpc_launch_count(const smpc_mpc *m);

/* ------------------------------------------------- multi-input MPC layer (BASELINE config 3) */
/* The reference class is single-input (N_C = N_O = 1, include/ModelPredictiveControlAPI.h:31-32).  For
 * multi-input plants (the 12-state / 4-input quadrotor of BASELINE config 3) this layer
 * builds the condensed QP the same way the reference does (setTransformations / setH / setFVars,
 * src/ModelPredictiveControlAPI.cpp:180-263,303-307: eliminate the states through powers of Ad, dense
 * Hessian 2(Su' Qbar Su + Rbar), gradient linear in the measured state and the reference) with
 *   decision  z = [u_0; ...; u_{N-1}]  (n = N*nu),   x_{k+1} = Ad x_k + Bd u_k,
 *   cost      sum_{k=1..N} (x_k - xr)' Q (x_k - xr) + sum_{k<N} u_k' R u_k      (Q, R diagonal),
 *   rows      [I; -I] z <= [umax; -umin] with l = -DBL_MAX: the two-sided input box written as 2n one-sided
 *             rows, as the reference writes its PWM limit (cpp:42,335)  ->  m = 2*N*nu.
 * Every controller shares the plant (shared-factor regime); per controller: x0 and xr (nx each). */
typedef struct smpc_mimo smpc_mimo;
typedef struct smpc_mimo_config {
  int horizon, nx, nu;
  const double *Ad;     /* nx*nx row-major */
  const double *Bd;     /* nx*nu row-major */
  const double *Q;      /* nx: diagonal state weight */
  const double *R;      /* nu: diagonal input weight */
  const double *umin;   /* nu */
  const double *umax;   /* nu */
} smpc_mimo_config;
/* assembles Su, H, Fx, Fr ON THE DEVICE, then sets up the batched shared-factor solver */
int smpc_mimo_create(smpc_mimo **out, int device, const smpc_mimo_config *cfg, int batch, const smpc_settings *settings);
/* JSON keys: "horizon", "Ad", "Bd", "Q" (nx diagonal), "R" (nu diagonal), "umin", "umax" */
int smpc_mimo_create_from_json(smpc_mimo **out, int device, const char *json_path, int batch, const smpc_settings *settings);
int smpc_mimo_destroy(smpc_mimo *m);
int smpc_mimo_set_stream(smpc_mimo *m, void *cuda_stream);
int smpc_mimo_dims(const smpc_mimo *m, int *horizon, int *nx, int *nu, int *n, int *mrows, int *batch);
smpc_solver *smpc_mimo_solver(smpc_mimo *m);
/* "H" (n*n) "A" (m*n) "ub" (m) "Fx" (n*nx) "Fr" (n*nx) "Su" (N*nx x n) "Sx" (N*nx x nx), row-major */
int smpc_mimo_get_matrix(smpc_mimo *m, const char *name, double *out, int capacity);
/* measured states x0:[batch][nx] and state references xr:[batch][nx] (held over the horizon); NULL keeps the value */
int smpc_mimo_set_state(smpc_mimo *m, const double *x0, const double *xr, int loc);
/* one controller step for every instance: q = Fx x0 + Fr xr -> updateGradient -> solve; asynchronous */
int smpc_mimo_controller_step(smpc_mimo *m);
/* the first move u_0 of every instance, u0:[batch][nu]: the solver's iterate for SOLVED / SOLVED_INACCURATE /
 * MAX_ITER_REACHED (check the status from smpc_solver_get_info before applying it), NaN for the infeasible statuses
 * (OSQP stores no solution there) */
int smpc_mimo_get_control(smpc_mimo *m, double *u0, int loc);
long long smpc_mimo_launch_count(const smpc_mimo *m);

/* ---------------------------------------------- wire format and asynchronous state feed (off the hot path) */
/* One frame of the reference's ASCII protocol, parsed as SerialPort::getDataFromSerial does (src/SerialPort.cpp:106-138):
 * five space-separated numbers "dt x1 x2 x3 x4" stored through float; frames of <= 30 bytes are rejected as readPort does
 * (cpp:146).  Returns 1 when the frame is accepted, 0 otherwise.  dt is returned (the reference loses it: readPort takes it
 * by value, cpp:142). */
int smpc_wire_parse_frame(const char *buf, int nbytes, double *dt, double *X /* [4] */);
/* The control as SerialPort::writePort formats it (std::to_string(U), cpp:165); max_chars = 8 reproduces the reference's
 * sizeof(char*) truncation, <= 0 sends the whole number.  Returns the number of characters written (NUL excluded). */
int smpc_wire_format_control(double U, char *out, int capacity, int max_chars);
/* A reader thread on file descriptor fd (a serial port, pipe or pty the caller opened and configured) keeps the most recent
 * accepted frame, so that one slow device never blocks the batch: smpc_feed_latest is non-blocking and returns 1 when a frame
 * newer than *seq is available (then *seq, dt, X[4] are updated), 0 otherwise.  A frame ends at '\n' or at 42 bytes. */
typedef struct smpc_feed smpc_feed;
int smpc_feed_open(smpc_feed **out, int fd);
int smpc_feed_latest(smpc_feed *f, long long *seq, double *dt, double *X /* [4] */);
int smpc_feed_stats(smpc_feed *f, long long *accepted, long long *rejected);
int smpc_feed_close(smpc_feed *f);   /* stops and joins the thread; does not close fd */

#ifdef __cplusplus
}
#endif
#endif
