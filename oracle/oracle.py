"""ctypes binding of the C oracle (oracle/osqp_port.c, oracle/mpc_assembly.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE -- see oracle/osqp_port.h for the parity
status ("parity unpinned" against real osqp-eigen; pinned by exact KKT known
answers).
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")

SOLVED, SOLVED_INACCURATE, MAX_ITER_REACHED = 1, 2, -2
PRIMAL_INFEASIBLE, DUAL_INFEASIBLE, UNSOLVED = -3, -4, -10
PRIMAL_INFEASIBLE_INACCURATE, DUAL_INFEASIBLE_INACCURATE = 3, 4


class Settings(C.Structure):
    _fields_ = [
        ("rho", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double),
        ("eps_prim_inf", C.c_double), ("eps_dual_inf", C.c_double),
        ("adaptive_rho_tolerance", C.c_double),
        ("max_iter", C.c_int), ("check_termination", C.c_int), ("scaling", C.c_int),
        ("adaptive_rho", C.c_int), ("adaptive_rho_interval", C.c_int),
        ("warm_start", C.c_int), ("scaled_termination", C.c_int),
        ("polish", C.c_int), ("polish_refine_iter", C.c_int), ("delta", C.c_double),
    ]


_SOURCES = ("osqp_port.c", "mpc_assembly.c", "batch_drivers.c", "osqp_port.h", "mpc_assembly.h")


def _stale(target):
    return not os.path.exists(target) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(target) for f in _SOURCES)


def build(force=False):
    """Compile the C oracle (gcc) into oracle/_build/; and oracle/_ref when the reference is present."""
    if force or _stale(_LIB):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


PERF_FLAGS = "-O3 -march=native -ffp-contract=fast -fassociative-math -fno-signed-zeros -fno-trapping-math -fno-math-errno"


def build_perf():
    """The same sources compiled for speed (PERF_FLAGS) ON this machine: used only by bench.py's CPU arm.  The file is
    named after the host's CPU flags so that a library built in one container is never loaded on another CPU."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    out = os.path.join(_HERE, "_build", "liboracle_perf_%s.so" % hashlib.sha1(flags.encode()).hexdigest()[:10])
    if _stale(out):
        subprocess.check_call(["make", "-C", _HERE, "-s", "perf", "PERF_OUT=" + out, "PERF_CFLAGS=" + PERF_FLAGS + " -fPIC -std=gnu11"])
    return out


_lib = None
_perf = None


def _bind(L):
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.orc_default_settings.argtypes = [C.POINTER(Settings)]
    L.orc_setup.restype = C.c_void_p
    L.orc_setup.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, C.POINTER(Settings)]
    L.orc_cleanup.argtypes = [C.c_void_p]
    for f in ("orc_update_lin_cost", "orc_update_lower_bound", "orc_update_upper_bound"):
        getattr(L, f).argtypes = [C.c_void_p, dp]
        getattr(L, f).restype = C.c_int
    L.orc_update_bounds.argtypes = [C.c_void_p, dp, dp]
    L.orc_warm_start.argtypes = [C.c_void_p, dp, dp]
    L.orc_cold_start.argtypes = [C.c_void_p]
    L.orc_reset.argtypes = [C.c_void_p]
    L.orc_solve.argtypes = [C.c_void_p]
    L.orc_get_solution.argtypes = [C.c_void_p, dp, dp]
    L.orc_get_info.argtypes = [C.c_void_p, dp]
    L.orc_get_scaling.argtypes = [C.c_void_p, dp, dp, dp]
    L.orc_get_status_polish.argtypes = [C.c_void_p]
    L.orc_get_scaled_data.argtypes = [C.c_void_p, dp, dp]
    L.orc_get_iterates.argtypes = [C.c_void_p, dp, dp, dp]
    L.orc_solve_batch.restype = C.c_double
    L.orc_solve_batch.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, C.POINTER(Settings), C.c_int,
                                  dp, dp, dp, C.c_int, dp, dp, ip, ip]
    L.orc_mpc_build.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double,
                                C.c_int, C.c_double] + [dp] * 11
    L.orc_mpc_step_vectors.argtypes = [C.c_int, C.c_int] + [dp] * 7 + [C.c_double, dp, dp, dp]
    L.orc_mpc_lower_bound.restype = C.c_double
    L.orc_plant_batch.restype = C.c_double
    L.orc_plant_batch.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                  C.c_double, dp, dp, dp, C.POINTER(Settings), C.c_int, dp, ip, ip]
    L.orc_closed_loop.restype = C.c_double
    L.orc_closed_loop.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                  C.c_double, dp, dp, C.c_double, C.c_int, ip, C.c_int, C.c_int, C.POINTER(Settings), C.c_int,
                                  C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), dp]
    return L


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = _bind(C.CDLL(_LIB))
    return _lib


def perf_lib():
    """liboracle built with PERF_FLAGS on this host (bench.py's CPU arm); the parity tests use lib()."""
    global _perf
    if _perf is None:
        _perf = _bind(C.CDLL(build_perf()))
    return _perf


def _p(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def default_settings(**kw):
    s = Settings()
    lib().orc_default_settings(C.byref(s))
    for k, v in kw.items():
        if not hasattr(s, k):
            raise KeyError(k)
        setattr(s, k, v)
    return s


class Solver:
    """One OSQP-equivalent solver instance (the CPU twin of OsqpEigen::Solver)."""

    def __init__(self, P, q, A, l, u, settings=None, **kw):
        P, A = _c(P), _c(A)
        self.n, self.m = P.shape[0], A.shape[0]
        self.settings = settings if settings is not None else default_settings(**kw)
        q, l, u = _c(q), _c(l), _c(u)
        self._h = lib().orc_setup(self.n, self.m, _p(P), _p(q), _p(A), _p(l), _p(u), C.byref(self.settings))
        if not self._h:
            raise ValueError("orc_setup failed (invalid data)")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_cleanup(self._h)
            self._h = None

    def update_lin_cost(self, q):
        return lib().orc_update_lin_cost(self._h, _p(_c(q)))

    def update_upper_bound(self, u):
        return lib().orc_update_upper_bound(self._h, _p(_c(u)))

    def update_lower_bound(self, l):
        return lib().orc_update_lower_bound(self._h, _p(_c(l)))

    def update_bounds(self, l, u):
        return lib().orc_update_bounds(self._h, _p(_c(l)), _p(_c(u)))

    def warm_start(self, x, y):
        return lib().orc_warm_start(self._h, _p(_c(x)), _p(_c(y)))

    def cold_start(self):
        lib().orc_cold_start(self._h)

    def reset(self):
        lib().orc_reset(self._h)

    def solve(self):
        rc = lib().orc_solve(self._h)
        x, y, info = np.empty(self.n), np.empty(self.m), np.empty(8)
        lib().orc_get_solution(self._h, _p(x), _p(y))
        lib().orc_get_info(self._h, _p(info))
        return dict(rc=rc, x=x, y=y, status=int(info[0]), iter=int(info[1]), rho_updates=int(info[2]),
                    status_polish=lib().orc_get_status_polish(self._h),
                    rho=info[3], obj=info[4], pri_res=info[5], dua_res=info[6], rho_estimate=info[7])

    def scaling(self):
        D, E, c = np.empty(self.n), np.empty(self.m), C.c_double()
        lib().orc_get_scaling(self._h, _p(D), _p(E), C.byref(c))
        return D, E, c.value

    def scaled_data(self):
        P, A = np.empty((self.n, self.n)), np.empty((self.m, self.n))
        lib().orc_get_scaled_data(self._h, _p(P), _p(A))
        return P, A

    def iterates(self):
        x, z, y = np.empty(self.n), np.empty(self.m), np.empty(self.m)
        lib().orc_get_iterates(self._h, _p(x), _p(z), _p(y))
        return x, z, y


def solve_batch(P, A, l0, u0, q, u, l=None, settings=None, nthreads=1, perf=False, **kw):
    """'One solver per core' batch of independent cold solves; returns dict with seconds."""
    P, A, q, u = _c(P), _c(A), _c(q), _c(u)
    l0, u0, l = _c(l0), _c(u0), _c(l)
    n, m, B = P.shape[0], A.shape[0], q.shape[0]
    s = settings if settings is not None else default_settings(**kw)
    x, y = np.empty((B, n)), np.empty((B, m))
    st, it = np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
    secs = (perf_lib() if perf else lib()).orc_solve_batch(n, m, _p(P), _p(A), _p(l0), _p(u0), C.byref(s), B, _p(q), _p(l), _p(u),
                                 nthreads, _p(x), _p(y), st.ctypes.data_as(C.POINTER(C.c_int)),
                                 it.ctypes.data_as(C.POINTER(C.c_int)))
    return dict(x=x, y=y, status=st, iter=it, seconds=secs)


# --------------------------------------------------------------------------- MPC assembly
def load_config(path):
    """Reads MPC_API.json (reference schema, cpp:16,19,113-116,138-140, plus additive keys)."""
    with open(path) as f:
        cfg = json.load(f)
    out = dict(
        Ad=np.array(cfg["Ad"], dtype=np.float64), Bd=np.array(cfg["Bd"], dtype=np.float64).reshape(-1),
        Cd=np.array(cfg["Cd"], dtype=np.float64).reshape(-1), K=np.array(cfg["K"], dtype=np.float64).reshape(-1),
        Q=float(np.array(cfg["Q"]).reshape(-1)[0]), R=float(np.array(cfg["R"]).reshape(-1)[0]),
        RD=float(np.array(cfg["RD"]).reshape(-1)[0]), xref=float(cfg["xref"]),
        N=int(cfg.get("horizon", 15)), n_state_rows=int(cfg.get("n_state_rows", 10)),
        u_limit=float(cfg.get("u_limit", 255.0)),
    )
    return out


def mpc_build(Ad, Bd, Cd, K, Q, R, RD, N=15, n_state_rows=10, u_limit=255.0, **_):
    Ad, Bd, Cd, K = _c(Ad), _c(Bd).reshape(-1), _c(Cd).reshape(-1), _c(K).reshape(-1)
    nx = Ad.shape[0]
    o = dict(H=np.empty((N, N)), Gbar=np.empty((2 * N, N)), Fx=np.empty((N, nx)), Fu=np.empty(N),
             Fr=np.empty((N, N)), Sbar=np.empty((2 * N, nx)), Ku=np.empty(2 * N), W0=np.empty(2 * N),
             Sx=np.empty((N, nx)), Su=np.empty((N, N)), CAB=np.empty(N))
    lib().orc_mpc_build(N, nx, _p(Ad), _p(Bd), _p(Cd), _p(K), Q, R, RD, n_state_rows, u_limit,
                        *[_p(o[k]) for k in ("H", "Gbar", "Fx", "Fu", "Fr", "Sbar", "Ku", "W0", "Sx", "Su", "CAB")])
    o["lb"] = np.full(2 * N, lib().orc_mpc_lower_bound())
    o["N"], o["nx"] = N, nx
    return o


def mpc_step_vectors(mats, X, U, ref):
    N, nx = mats["N"], mats["nx"]
    X = _c(X).reshape(-1)
    ref = _c(np.broadcast_to(np.asarray(ref, dtype=np.float64), (N,)))
    f, ub = np.empty(N), np.empty(2 * N)
    lib().orc_mpc_step_vectors(N, nx, _p(mats["Fx"]), _p(mats["Fu"]), _p(mats["Fr"]), _p(mats["Sbar"]),
                               _p(mats["Ku"]), _p(mats["W0"]), _p(X), float(U), _p(ref), _p(f), _p(ub))
    return f, ub


def mpc_batch_vectors(mats, X, U, ref):
    """Vectorised f / ub for a batch (numpy; same formulas as orc_mpc_step_vectors)."""
    X, U, ref = np.atleast_2d(X), np.asarray(U, dtype=np.float64).reshape(-1), np.asarray(ref, dtype=np.float64)
    N = mats["N"]
    if ref.ndim == 1:
        ref = np.repeat(ref[:, None], N, axis=1)
    f = (X @ mats["Fx"].T + U[:, None] * mats["Fu"][None, :]) + ref @ mats["Fr"].T
    ub = (mats["W0"][None, :] + X @ mats["Sbar"].T) + U[:, None] * mats["Ku"][None, :]
    return f, ub


def plant_batch(cfg, Ad, Bd, X, U, ref, N, settings=None, nthreads=1, perf=False, **kw):
    """One controller PER PLANT (BASELINE config 4): constructor (builders + solver setup) and one controllerStep for every
    instance, one controller per core.  Ad [B][nx][nx], Bd [B][nx].  Returns dict with x, status, iter, seconds."""
    Ad, Bd, X, U, ref = _c(Ad), _c(Bd), _c(X), _c(U), _c(ref)
    B, nx = Ad.shape[0], Ad.shape[1]
    s = settings if settings is not None else default_settings(**kw)
    x = np.empty((B, N))
    st, it = np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
    ip = C.POINTER(C.c_int)
    secs = (perf_lib() if perf else lib()).orc_plant_batch(
        N, nx, B, _p(Ad), _p(Bd), _p(_c(cfg["Cd"])), _p(_c(cfg["K"])), cfg["Q"], cfg["R"], cfg["RD"],
        int(cfg.get("n_state_rows", 10)), float(cfg.get("u_limit", 255.0)), _p(X), _p(U), _p(ref), C.byref(s), nthreads,
        _p(x), st.ctypes.data_as(ip), it.ctypes.data_as(ip))
    return dict(x=x, status=st, iter=it, seconds=secs)


def closed_loop(cfg, X, U, N, steps, amp, period=0, phase=None, skip=0, settings=None, nthreads=1, perf=False, **kw):
    """Warm-started closed loop (the reference's main loop, src/solver.cpp:43-74, with a synthetic plant) of B controllers
    sharing the plant, one controller per core.  Returns dict with X, U, not_solved, iterations, seconds (timed steps =
    steps - skip), step_seconds (thread 0)."""
    X, U = _c(X).copy(), _c(U).copy()
    B, nx = X.shape
    s = settings if settings is not None else default_settings(**kw)
    ph = None if phase is None else np.ascontiguousarray(phase, dtype=np.int32)
    bad, it = C.c_longlong(), C.c_longlong()
    step_secs = np.zeros(steps)
    secs = (perf_lib() if perf else lib()).orc_closed_loop(
        N, nx, B, _p(_c(cfg["Ad"])), _p(_c(cfg["Bd"]).reshape(-1)), _p(_c(cfg["Cd"])), _p(_c(cfg["K"])), cfg["Q"], cfg["R"],
        cfg["RD"], int(cfg.get("n_state_rows", 10)), float(cfg.get("u_limit", 255.0)), _p(X), _p(U), float(amp), int(period),
        None if ph is None else ph.ctypes.data_as(C.POINTER(C.c_int)), int(steps), int(skip), C.byref(s), nthreads,
        C.byref(bad), C.byref(it), _p(step_secs))
    return dict(X=X, U=U, not_solved=bad.value, iterations=it.value, seconds=secs, step_seconds=step_secs)


# --------------------------------------------------------------------------- exact KKT (solver independent)
def exact_qp_active_set(P, q, A, u, active):
    """Solves the equality-constrained QP  min .5x'Px+q'x  s.t. A[active] x = u[active]  exactly
    (dense KKT solve) and returns (x, y) with y zero off the active set.  Solver independent."""
    P, q, A, u = map(lambda a: np.asarray(a, dtype=np.float64), (P, q, A, u))
    n, m = P.shape[0], A.shape[0]
    act = np.asarray(sorted(active), dtype=int)
    k = len(act)
    KKT = np.zeros((n + k, n + k))
    KKT[:n, :n] = P
    KKT[:n, n:] = A[act].T
    KKT[n:, :n] = A[act]
    rhs = np.concatenate([-q, u[act]])
    sol = np.linalg.solve(KKT, rhs)
    y = np.zeros(m)
    y[act] = sol[n:]
    return sol[:n], y


def kkt_report(P, q, A, l, u, x, y):
    """Solver-independent optimality measures of (x, y) for  min .5x'Px+q'x, l<=Ax<=u."""
    Ax = A @ x
    stat = np.abs(P @ x + q + A.T @ y).max()
    lo = np.where(np.isfinite(l) & (l > -1e300), l - Ax, -np.inf)
    feas = max(0.0, (Ax - u).max(), lo.max())
    yp, ym = np.maximum(y, 0), np.minimum(y, 0)
    comp = max(np.abs(yp * (u - Ax)).max(), np.abs(np.where(np.isfinite(lo), ym * (Ax - np.where(l > -1e300, l, 0)), 0)).max())
    return dict(stationarity=stat, infeasibility=feas, complementarity=comp)


# --------------------------------------------------------------------------- multi-input condensed MPC (config 3)
def load_mimo_config(path):
    """config/quadrotor.json: horizon, Ad, Bd, Q (diagonal), R (diagonal), umin, umax."""
    with open(path) as f:
        cfg = json.load(f)
    return dict(Ad=np.array(cfg["Ad"], dtype=np.float64), Bd=np.array(cfg["Bd"], dtype=np.float64),
                Q=np.array(cfg["Q"], dtype=np.float64).reshape(-1), R=np.array(cfg["R"], dtype=np.float64).reshape(-1),
                umin=np.array(cfg["umin"], dtype=np.float64).reshape(-1), umax=np.array(cfg["umax"], dtype=np.float64).reshape(-1),
                N=int(cfg["horizon"]))


def mimo_build(Ad, Bd, Q, R, umin, umax, N, **_):
    """Condensed QP of the multi-input MPC layer, built the way the reference builds its own
    (setTransformations cpp:180-208: states eliminated through powers of Ad; setH cpp:247-263:
    H = 2(Su'Qbar Su + Rbar); setFVars cpp:303-307: gradient maps; two-sided limit as 2n one-sided
    rows with l = -DBL_MAX, cpp:42,335), with nu inputs and full-state tracking (include/solvempc_b200.h).
    Plain numpy with explicit Su / Sx -- deliberately a different evaluation order than the device kernel."""
    Ad, Bd = _c(Ad), _c(Bd)
    nx, nu = Bd.shape
    n = N * nu
    Sx = np.zeros((N * nx, nx))
    Su = np.zeros((N * nx, n))
    Ak = np.eye(nx)
    AB = []
    for k in range(N):
        AB.append(Ak @ Bd)          # Ad^k Bd
        Ak = Ak @ Ad
        Sx[k * nx:(k + 1) * nx] = Ak   # x_{k+1} row block = Ad^(k+1)
    for k in range(1, N + 1):
        for j in range(k):
            Su[(k - 1) * nx:k * nx, j * nu:(j + 1) * nu] = AB[k - 1 - j]
    Qbar = np.tile(np.asarray(Q, dtype=np.float64), N)
    Rbar = np.tile(np.asarray(R, dtype=np.float64), N)
    SQ = Su.T * Qbar[None, :]
    H = 2.0 * (SQ @ Su + np.diag(Rbar))
    Fx = 2.0 * (SQ @ Sx)
    Fr = -2.0 * (SQ @ np.tile(np.eye(nx), (N, 1)))
    A = np.vstack([np.eye(n), -np.eye(n)])
    ub = np.concatenate([np.tile(umax, N), -np.tile(umin, N)])
    lb = np.full(2 * n, lib().orc_mpc_lower_bound())
    return dict(H=H, A=A, ub=ub, lb=lb, Fx=Fx, Fr=Fr, Su=Su, Sx=Sx, N=N, nx=nx, nu=nu)


def mimo_batch_vectors(mats, x0, xr):
    """q = Fx x0 + Fr xr for a batch."""
    return np.atleast_2d(x0) @ mats["Fx"].T + np.atleast_2d(xr) @ mats["Fr"].T
