"""Host-side mirror of the two reference-facing interfaces, on top of the C ABI.

``BatchedSolver``  ~ OsqpEigen::Solver (reference include/ModelPredictiveControlAPI.h:144) for a batch of QPs
``BatchedModelPredictiveControlAPI`` ~ class ModelPredictiveControlAPI (reference h:47-243) for a batch of controllers

Buffers may be numpy arrays (host) or torch CUDA tensors (device); torch is used only for device memory and
streams.  All arithmetic happens inside libsolvempc_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib as L


def _loc_ptr(a, count, name):
    """(pointer, loc, keepalive) for a float64 numpy array or torch tensor with `count` elements."""
    if a is None:
        return None, L.HOST, None
    if isinstance(a, np.ndarray) or not hasattr(a, "data_ptr"):
        arr = np.ascontiguousarray(a, dtype=np.float64)
        if arr.size != count:
            raise ValueError(f"{name}: expected {count} elements, got {arr.size}")
        return arr.ctypes.data, L.HOST, arr
    import torch
    if a.dtype != torch.float64 or not a.is_contiguous() or a.numel() != count:
        raise ValueError(f"{name}: expected a contiguous float64 tensor with {count} elements")
    return a.data_ptr(), (L.DEVICE if a.is_cuda else L.HOST), a


def _np_ptr(a):
    return None if a is None else a.ctypes.data


class BatchedSolver:
    """A batch of OSQP-equivalent solvers sharing P and A (shared-factor regime)."""

    def __init__(self, P, A, l0=None, u0=None, batch=1, device=0, settings=None, q0=None, _handle=None, **kw):
        self._own = _handle is None
        self.settings = settings if settings is not None else L.default_settings(**kw)
        if _handle is not None:
            self._h = C.c_void_p(_handle)
        else:
            P = np.ascontiguousarray(P, dtype=np.float64)
            A = np.ascontiguousarray(A, dtype=np.float64)
            n, m = P.shape[0], A.shape[0]
            if P.shape != (n, n) or (m and A.shape != (m, n)):
                raise ValueError("P must be n x n and A m x n")
            l0 = None if l0 is None else np.ascontiguousarray(l0, dtype=np.float64)
            u0 = None if u0 is None else np.ascontiguousarray(u0, dtype=np.float64)
            q0 = None if q0 is None else np.ascontiguousarray(q0, dtype=np.float64)
            h = C.c_void_p()
            L.check(L.lib().smpc_solver_create_shared(C.byref(h), device, n, m, batch, P.ctypes.data, A.ctypes.data,
                                                      _np_ptr(q0), _np_ptr(l0), _np_ptr(u0), C.byref(self.settings)))
            self._h = h
        n, m, b = C.c_int(), C.c_int(), C.c_int()
        L.check(L.lib().smpc_solver_dims(self._h, C.byref(n), C.byref(m), C.byref(b)))
        self.n, self.m, self.batch = n.value, m.value, b.value

    @classmethod
    def batched(cls, P, A, l0=None, u0=None, device=0, settings=None, **kw):
        """Per-instance regime: P [B][n][n] and A [B][m][n] differ per instance (numpy or CUDA tensors)."""
        st = settings if settings is not None else L.default_settings(**kw)
        B, n = P.shape[0], P.shape[1]
        m = A.shape[1]
        pp, lp, _k1 = _loc_ptr(P, B * n * n, "P")
        pa, la, _k2 = _loc_ptr(A, B * m * n, "A") if m else (None, lp, None)
        if m and lp != la:
            raise ValueError("P and A must live in the same place")
        l0 = None if l0 is None else np.ascontiguousarray(l0, dtype=np.float64)
        u0 = None if u0 is None else np.ascontiguousarray(u0, dtype=np.float64)
        h = C.c_void_p()
        L.check(L.lib().smpc_solver_create_batched(C.byref(h), device, n, m, B, pp, pa, lp, _np_ptr(l0), _np_ptr(u0), C.byref(st)))
        obj = cls(None, None, _handle=h.value, settings=st)
        obj._own = True
        return obj

    def close(self):
        if getattr(self, "_h", None) and self._own:
            try:
                L.lib().smpc_solver_destroy(self._h)
            except TypeError:      # interpreter shutdown: the module globals are already gone
                pass
        self._h = None

    __del__ = close

    def set_stream(self, cuda_stream):
        L.check(L.lib().smpc_solver_set_stream(self._h, C.c_void_p(cuda_stream)))

    # --- updates (osqp-eigen: updateGradient / updateUpperBound / updateLowerBound / updateBounds)
    def update_gradient(self, q):
        p, loc, _k = _loc_ptr(q, self.batch * self.n, "q")
        L.check(L.lib().smpc_solver_update_lin_cost(self._h, p, loc))

    def update_upper_bound(self, u):
        p, loc, _k = _loc_ptr(u, self.batch * self.m, "u")
        L.check(L.lib().smpc_solver_update_upper_bound(self._h, p, loc))

    def update_lower_bound(self, l):
        p, loc, _k = _loc_ptr(l, self.batch * self.m, "l")
        L.check(L.lib().smpc_solver_update_lower_bound(self._h, p, loc))

    def update_bounds(self, l, u):
        self.update_lower_bound(l)
        self.update_upper_bound(u)

    def warm_start(self, x, y):
        px, loc, _k1 = _loc_ptr(x, self.batch * self.n, "x")
        py, loc2, _k2 = _loc_ptr(y, self.batch * self.m, "y")
        if loc != loc2:
            raise ValueError("x and y must live in the same place")
        L.check(L.lib().smpc_solver_warm_start(self._h, px, py, loc))

    def cold_start(self):
        L.check(L.lib().smpc_solver_cold_start(self._h))

    def reset(self):
        L.check(L.lib().smpc_solver_reset(self._h))

    def set_cold_solves(self, on=True):
        """Every solve behaves like a freshly constructed solver (x = z = y = 0, rho = settings.rho)."""
        L.check(L.lib().smpc_solver_set_cold_solves(self._h, int(on)))

    def set_scheduling(self, on=True):
        """Longest-expected-first ordering of the small kernel's work queue (default on; results identical)."""
        L.check(L.lib().smpc_solver_set_scheduling(self._h, int(on)))

    def enable_timing(self, on=True):
        L.check(L.lib().smpc_solver_enable_timing(self._h, int(on)))

    def kernel_ms(self, reset=True):
        """(total ms, launches) of the ADMM kernel since the last reset, from CUDA events on its stream."""
        ms, cnt = C.c_double(), C.c_int()
        L.check(L.lib().smpc_solver_kernel_ms(self._h, C.byref(ms), C.byref(cnt), int(reset)))
        return ms.value, cnt.value

    def set_polish(self, on=True, delta=1e-6, refine_iter=3):
        """settings()->setPolish / setDelta / setPolishRefineIter (OSQP polish.c); off by default as in the reference."""
        L.check(L.lib().smpc_solver_set_polish(self._h, int(on), float(delta), int(refine_iter)))

    def polish_status(self):
        """info->status_polish per instance: 1 polished, -1 unsuccessful, 0 not run."""
        st = np.empty(self.batch, np.int32)
        L.check(L.lib().smpc_solver_get_polish_status(self._h, st.ctypes.data, L.HOST))
        return st

    def solve(self):
        """Asynchronous on the handle's stream (osqp-eigen solve())."""
        L.check(L.lib().smpc_solver_solve(self._h))

    def sync(self):
        L.check(L.lib().smpc_solver_sync(self._h))

    def solution(self, want_y=True):
        x = np.empty((self.batch, self.n))
        y = np.empty((self.batch, self.m)) if want_y else None
        L.check(L.lib().smpc_solver_get_solution(self._h, x.ctypes.data, _np_ptr(y), L.HOST))
        return (x, y) if want_y else x

    def solution_into(self, x=None, y=None):
        """Copies into caller buffers (numpy / pinned torch host tensors / CUDA tensors)."""
        px, lx, _k1 = _loc_ptr(x, self.batch * self.n, "x")
        py, ly, _k2 = _loc_ptr(y, self.batch * self.m, "y")
        loc = lx if x is not None else ly
        L.check(L.lib().smpc_solver_get_solution(self._h, px, py, loc))

    def info(self):
        B = self.batch
        st, it, ru = np.empty(B, np.int32), np.empty(B, np.int32), np.empty(B, np.int32)
        obj, pr, du, rho = np.empty(B), np.empty(B), np.empty(B), np.empty(B)
        L.check(L.lib().smpc_solver_get_info(self._h, st.ctypes.data, it.ctypes.data, obj.ctypes.data, pr.ctypes.data,
                                             du.ctypes.data, rho.ctypes.data, ru.ctypes.data, L.HOST))
        return dict(status=st, iter=it, obj=obj, pri_res=pr, dua_res=du, rho=rho, rho_updates=ru)

    def status_into(self, status):
        """status: int32 numpy / torch tensor of length batch."""
        if hasattr(status, "data_ptr"):
            p, loc = status.data_ptr(), (L.DEVICE if status.is_cuda else L.HOST)
        else:
            p, loc = status.ctypes.data, L.HOST
        L.check(L.lib().smpc_solver_get_info(self._h, p, None, None, None, None, None, None, loc))

    def count_solved(self):
        c = C.c_int()
        L.check(L.lib().smpc_solver_count_solved(self._h, C.byref(c)))
        return c.value

    def scaling(self):
        D, E, c = np.empty(self.n), np.empty(self.m), C.c_double()
        L.check(L.lib().smpc_solver_get_scaling(self._h, D.ctypes.data, E.ctypes.data, C.byref(c)))
        return D, E, c.value

    @property
    def launches(self):
        return L.lib().smpc_solver_launch_count(self._h)

    @property
    def kernel_name(self):
        return L.lib().smpc_solver_kernel_name(self._h).decode()

    @property
    def row_pairs(self):
        """m / 2 when the rows come as [G; -G] pairs and the kernel runs its iteration GEMMs on the top half only."""
        return L.lib().smpc_solver_row_pairs(self._h)


def shared_plan_inspect(P, A, l0=None, u0=None, settings=None, q0=None, **kw):
    """Host-only view of the shared-factor plan (no device needed)."""
    s = settings if settings is not None else L.default_settings(**kw)
    P = np.ascontiguousarray(P, dtype=np.float64)
    A = np.ascontiguousarray(A, dtype=np.float64)
    n, m = P.shape[0], A.shape[0]
    l0 = None if l0 is None else np.ascontiguousarray(l0, dtype=np.float64)
    u0 = None if u0 is None else np.ascontiguousarray(u0, dtype=np.float64)
    q0 = None if q0 is None else np.ascontiguousarray(q0, dtype=np.float64)
    o = dict(D=np.empty(n), E=np.empty(m), lam=np.empty(n), V=np.empty((n, n)), SG=np.empty((n, n)),
             W=np.empty((m, n)), PVT=np.empty((n, n)), VinvT=np.empty((n, n)), ctype=np.empty(m, np.int8))
    c = C.c_double()
    L.check(L.lib().smpc_shared_plan_inspect(n, m, P.ctypes.data, A.ctypes.data, _np_ptr(q0), _np_ptr(l0), _np_ptr(u0), C.byref(s),
                                             o["D"].ctypes.data, o["E"].ctypes.data, C.byref(c), o["lam"].ctypes.data,
                                             o["V"].ctypes.data, o["SG"].ctypes.data, o["W"].ctypes.data,
                                             o["PVT"].ctypes.data, o["VinvT"].ctypes.data, o["ctype"].ctypes.data))
    o["c"] = c.value
    return o


class BatchedModelPredictiveControlAPI:
    """A batch of the reference's controllers.  Member and method names follow the reference class
    (X, U, xref, controllerStep, H, Gbar, Fx, Fu, Fr, Sbar, Ku, W0, solver)."""

    MATRIX_SHAPES = {"H": ("N", "N"), "Gbar": ("2N", "N"), "Fx": ("N", "nx"), "Fu": ("N",), "Fr": ("N", "N"),
                     "Sbar": ("2N", "nx"), "Ku": ("2N",), "W0": ("2N",), "Sx": ("N", "nx"), "Su": ("N", "N"), "CAB": ("N",)}

    def __init__(self, config="./config/MPC_API.json", batch=1, device=0, settings=None, verbose=False, **kw):
        self.verbose = verbose
        self.settings = settings if settings is not None else L.default_settings(**kw)
        h = C.c_void_p()
        if isinstance(config, (str, bytes)):
            path = config if isinstance(config, bytes) else config.encode()
            L.check(L.lib().smpc_mpc_create_from_json(C.byref(h), device, path, batch, C.byref(self.settings)))
        else:
            cfg = L.MpcConfig()
            self._keep = [np.ascontiguousarray(config[k], dtype=np.float64) for k in ("Ad", "Bd", "Cd", "K")]
            cfg.horizon = int(config.get("horizon", config.get("N", 15)))
            cfg.nx = self._keep[0].shape[-1]
            cfg.n_state_rows = int(config.get("n_state_rows", 10))
            cfg.Q, cfg.R, cfg.RD = float(config["Q"]), float(config["R"]), float(config["RD"])
            cfg.u_limit, cfg.xref = float(config.get("u_limit", 255.0)), float(config.get("xref", 0.0))
            cfg.Ad, cfg.Bd, cfg.Cd, cfg.K = [a.ctypes.data for a in self._keep]
            cfg.per_instance = int(config.get("per_instance", 0))
            L.check(L.lib().smpc_mpc_create(C.byref(h), device, C.byref(cfg), batch, C.byref(self.settings)))
        self._h = h
        d = [C.c_int() for _ in range(5)]
        L.check(L.lib().smpc_mpc_dims(self._h, *[C.byref(v) for v in d]))
        self.mpcWindow, self.N_S, self.n_variables, self.n_constraints, self.batch = [v.value for v in d]
        self.solver = BatchedSolver(None, None, _handle=L.lib().smpc_mpc_solver(self._h), settings=self.settings)
        self.solverFlag = True

    def close(self):
        if getattr(self, "_h", None):
            try:
                L.lib().smpc_mpc_destroy(self._h)
            except TypeError:      # interpreter shutdown: the module globals are already gone
                pass
        self._h = None

    __del__ = close

    def set_stream(self, cuda_stream):
        L.check(L.lib().smpc_mpc_set_stream(self._h, C.c_void_p(cuda_stream)))

    def matrix(self, name, index=0):
        dims = {"N": self.mpcWindow, "2N": 2 * self.mpcWindow, "nx": self.N_S}
        shape = tuple(dims[s] for s in self.MATRIX_SHAPES[name])
        out = np.empty(shape)
        L.check(L.lib().smpc_mpc_get_matrix(self._h, name.encode(), index, out.ctypes.data, out.size))
        return out

    def set_state(self, X=None, U=None, ref=None):
        """Writes the public members X (h:186), U (h:187) and the reference held over the horizon (cpp:378-381)."""
        px, l1, _k1 = _loc_ptr(X, self.batch * self.N_S, "X")
        pu, l2, _k2 = _loc_ptr(U, self.batch, "U")
        pr, l3, _k3 = _loc_ptr(ref, self.batch, "ref")
        locs = {l for a, l in ((X, l1), (U, l2), (ref, l3)) if a is not None}
        if len(locs) > 1:
            raise ValueError("X, U, ref must live in the same place")
        L.check(L.lib().smpc_mpc_set_state(self._h, px, pu, pr, locs.pop() if locs else L.HOST))

    def controllerStep(self):
        """cpp:81-108 for every controller; returns True when every instance reports SOLVED (the reference
        returns solver.solve()'s flag)."""
        L.check(L.lib().smpc_mpc_controller_step(self._h))
        return self.solver.count_solved() == self.batch

    def controller_step_async(self):
        L.check(L.lib().smpc_mpc_controller_step(self._h))

    def controller_step_from(self, X, U, ref):
        """set_state(X, U, ref) + controllerStep in one asynchronous call (the body of the reference's loop,
        src/solver.cpp:45-55): device tensors and pinned host tensors are read where they lie by the step's first kernel."""
        px, l1, _k1 = _loc_ptr(X, self.batch * self.N_S, "X")
        pu, l2, _k2 = _loc_ptr(U, self.batch, "U")
        pr, l3, _k3 = _loc_ptr(ref, self.batch, "ref")
        if X is None or U is None or ref is None:
            raise ValueError("X, U and ref are all required")
        if len({l1, l2, l3}) > 1:
            raise ValueError("X, U, ref must live in the same place")
        L.check(L.lib().smpc_mpc_controller_step_from(self._h, px, pu, pr, l1))

    def plant_step(self):
        L.check(L.lib().smpc_mpc_plant_step(self._h))

    def closed_loop(self, steps, ref_amplitude=0.0, ref_period=0, phase=None, use_graph=True):
        """The reference's main loop (src/solver.cpp:43-74) for the whole batch on the device: `steps` times
        [square-wave reference ->] controllerStep -> synthetic plant step.  Returns (not_solved, iterations)."""
        ph = None if phase is None else np.ascontiguousarray(phase, dtype=np.int32)
        if ph is not None and ph.size != self.batch:
            raise ValueError("phase: expected one int per instance")
        bad, it = C.c_longlong(), C.c_longlong()
        L.check(L.lib().smpc_mpc_closed_loop(self._h, int(steps), float(ref_amplitude), int(ref_period),
                                             None if ph is None else ph.ctypes.data, int(use_graph), C.byref(bad), C.byref(it)))
        return bad.value, it.value

    def state(self):
        X, U = np.empty((self.batch, self.N_S)), np.empty(self.batch)
        L.check(L.lib().smpc_mpc_get_state(self._h, X.ctypes.data, U.ctypes.data, L.HOST))
        return X, U

    def control_into(self, U):
        p, loc, _k = _loc_ptr(U, self.batch, "U")
        L.check(L.lib().smpc_mpc_get_state(self._h, None, p, loc))

    def results_into(self, U, status):
        """U (float64) and the per-instance status (int32) of the last controllerStep in one transfer + one sync."""
        pu, l1, _k1 = _loc_ptr(U, self.batch, "U")
        if hasattr(status, "data_ptr"):
            ps, l2 = status.data_ptr(), (L.DEVICE if status.is_cuda else L.HOST)
        else:
            ps, l2 = status.ctypes.data, L.HOST
        if l1 != l2:
            raise ValueError("U and status must live in the same place")
        L.check(L.lib().smpc_mpc_get_control_status(self._h, pu, ps, l1))

    def bind_results(self, U=None, status=None):
        """Every later controllerStep also writes U (float64) and the statuses (int32) to these device or PINNED host tensors
        (what the reference's loop reads after the step, src/solver.cpp:55-60); valid after sync().  None, None unbinds."""
        self._bound = (U, status)                       # keep the buffers alive
        if U is None and status is None:
            L.check(L.lib().smpc_mpc_bind_results(self._h, None, None, L.HOST))
            return
        pu, l1, _k1 = _loc_ptr(U, self.batch, "U") if U is not None else (None, None, None)
        ps, l2 = None, None
        if status is not None:
            if hasattr(status, "data_ptr"):
                ps, l2 = status.data_ptr(), (L.DEVICE if status.is_cuda else L.HOST)
            else:
                ps, l2 = status.ctypes.data, L.HOST
        locs = {l for l in (l1, l2) if l is not None}
        if len(locs) > 1:
            raise ValueError("U and status must live in the same place")
        L.check(L.lib().smpc_mpc_bind_results(self._h, pu, ps, locs.pop()))

    def sync(self):
        L.check(L.lib().smpc_mpc_sync(self._h))

    def step_vectors(self):
        f, ub = np.empty((self.batch, self.n_variables)), np.empty((self.batch, self.n_constraints))
        L.check(L.lib().smpc_mpc_get_step_vectors(self._h, f.ctypes.data, ub.ctypes.data, L.HOST))
        return f, ub

    @property
    def launches(self):
        return L.lib().smpc_mpc_launch_count(self._h)


class BatchedMimoMPC:
    """A batch of multi-input condensed MPC controllers sharing one plant (BASELINE config 3: the 12-state /
    4-input quadrotor).  Same construction as the reference builders (cpp:180-263,303-307) generalised to nu
    inputs; see include/solvempc_b200.h (smpc_mimo_*)."""

    def __init__(self, config, batch=1, device=0, settings=None, **kw):
        self.settings = settings if settings is not None else L.default_settings(**kw)
        h = C.c_void_p()
        if isinstance(config, (str, bytes)):
            path = config if isinstance(config, bytes) else config.encode()
            L.check(L.lib().smpc_mimo_create_from_json(C.byref(h), device, path, batch, C.byref(self.settings)))
        else:
            cfg = L.MimoConfig()
            Ad = np.ascontiguousarray(config["Ad"], dtype=np.float64)
            Bd = np.ascontiguousarray(config["Bd"], dtype=np.float64).reshape(Ad.shape[0], -1)
            self._keep = [Ad, Bd] + [np.ascontiguousarray(config[k], dtype=np.float64).reshape(-1) for k in ("Q", "R", "umin", "umax")]
            cfg.horizon, cfg.nx, cfg.nu = int(config.get("horizon", config.get("N"))), Ad.shape[0], Bd.shape[1]
            cfg.Ad, cfg.Bd, cfg.Q, cfg.R, cfg.umin, cfg.umax = [a.ctypes.data for a in self._keep]
            L.check(L.lib().smpc_mimo_create(C.byref(h), device, C.byref(cfg), batch, C.byref(self.settings)))
        self._h = h
        d = [C.c_int() for _ in range(6)]
        L.check(L.lib().smpc_mimo_dims(self._h, *[C.byref(v) for v in d]))
        self.horizon, self.nx, self.nu, self.n_variables, self.n_constraints, self.batch = [v.value for v in d]
        self.solver = BatchedSolver(None, None, _handle=L.lib().smpc_mimo_solver(self._h), settings=self.settings)

    def close(self):
        if getattr(self, "_h", None):
            try:
                L.lib().smpc_mimo_destroy(self._h)
            except TypeError:      # interpreter shutdown: the module globals are already gone
                pass
        self._h = None

    __del__ = close

    def set_stream(self, cuda_stream):
        L.check(L.lib().smpc_mimo_set_stream(self._h, C.c_void_p(cuda_stream)))

    def matrix(self, name):
        N, nx, n, m = self.horizon, self.nx, self.n_variables, self.n_constraints
        shape = {"H": (n, n), "A": (m, n), "ub": (m,), "Fx": (n, nx), "Fr": (n, nx), "Su": (N * nx, n), "Sx": (N * nx, nx)}[name]
        out = np.empty(shape)
        L.check(L.lib().smpc_mimo_get_matrix(self._h, name.encode(), out.ctypes.data, out.size))
        return out

    def set_state(self, x0=None, xr=None):
        p0, l0, _k0 = _loc_ptr(x0, self.batch * self.nx, "x0")
        p1, l1, _k1 = _loc_ptr(xr, self.batch * self.nx, "xr")
        locs = {l for a, l in ((x0, l0), (xr, l1)) if a is not None}
        if len(locs) > 1:
            raise ValueError("x0 and xr must live in the same place")
        L.check(L.lib().smpc_mimo_set_state(self._h, p0, p1, locs.pop() if locs else L.HOST))

    def controllerStep(self):
        L.check(L.lib().smpc_mimo_controller_step(self._h))
        return self.solver.count_solved() == self.batch

    def controller_step_async(self):
        L.check(L.lib().smpc_mimo_controller_step(self._h))

    def control(self):
        u0 = np.empty((self.batch, self.nu))
        L.check(L.lib().smpc_mimo_get_control(self._h, u0.ctypes.data, L.HOST))
        return u0

    def control_into(self, u0):
        p, loc, _k = _loc_ptr(u0, self.batch * self.nu, "u0")
        L.check(L.lib().smpc_mimo_get_control(self._h, p, loc))

    @property
    def launches(self):
        return L.lib().smpc_mimo_launch_count(self._h)
