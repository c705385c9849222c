"""Solver-INDEPENDENT pins of the GPU results (no oracle/ in this file).

The CPU oracle was written from the same description of OSQP as the kernels, so agreement with it proves the two agree
with each other.  These tests check the GPU answers against facts that do not depend on any ADMM implementation:

  * SOLVED           -> the exact KKT system of the active set the GPU reports (dense numpy solve) reproduces the solution,
                        is primal feasible and has correctly signed duals: a KKT point of a convex QP is its global optimum;
  * PRIMAL_INFEASIBLE-> a phase-1 LP (scipy / HiGHS) proves {x : l <= A x <= u} empty, and a Farkas certificate exists;
  * DUAL_INFEASIBLE  -> an LP proves a direction d with P d = 0, q'd < 0, A d in the recession cone of [l, u] exists;
  * the reference's own, byte-identical config/MPC_API.json (tests/golden/MPC_API.reference.json, with its dead keys)
    parses through smpc_mpc_create_from_json and gives the golden matrices of the reference's compiled .cpp.
"""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, assume, given, settings
from hypothesis import strategies as hst
from scipy.optimize import linprog

import solvempc_b200 as sm
from problems import random_qp

pytestmark = pytest.mark.gpu
EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
INF = 1e20


def kkt_from_active_set(P, q, A, l, u, x, y, thr=1e-7):
    """Exact solution of the equality-constrained QP on the active set read off (x, y): rows with y > 0 sit at u, rows with
    y < 0 at l.  Returns (x*, y*, active mask)."""
    n, m = P.shape[0], A.shape[0]
    scale = max(np.abs(y).max(), 1e-12)
    up, lo = y > thr * scale, y < -thr * scale
    act = np.nonzero(up | lo)[0]
    rhs_b = np.where(up, u, l)[act]
    k = len(act)
    K = np.zeros((n + k, n + k))
    K[:n, :n] = P
    K[:n, n:] = A[act].T
    K[n:, :n] = A[act]
    sol = np.linalg.lstsq(K, np.concatenate([-q, rhs_b]), rcond=None)[0]
    ys = np.zeros(m)
    ys[act] = sol[n:]
    return sol[:n], ys, up, lo


def assert_kkt_optimal(P, q, A, l, u, x, y, rel=1e-4):
    xs, ys, up, lo = kkt_from_active_set(P, q, A, l, u, x, y)
    Ax = A @ xs
    tol = 1e-6 * max(1.0, np.abs(Ax).max())
    assert (Ax <= u + tol).all() and (Ax >= l - tol).all(), "KKT point of the reported active set is not primal feasible"
    assert (ys[up] >= -1e-9).all() and (ys[lo] <= 1e-9).all(), "dual sign wrong on the reported active set"
    assert np.abs(P @ xs + q + A.T @ ys).max() < 1e-8 * max(1.0, np.abs(q).max())
    assert np.abs(x - xs).max() <= rel * max(np.abs(xs).max(), 1e-3), "GPU solution is not the optimum within 1e-4 relative"


def lp_set_is_empty(A, l, u):
    """phase-1 LP: is {x : l <= A x <= u} empty?  (HiGHS; independent of any QP solver)"""
    n = A.shape[1]
    fin_u, fin_l = u < INF, l > -INF
    Aub = np.vstack([A[fin_u], -A[fin_l]])
    bub = np.concatenate([u[fin_u], -l[fin_l]])
    r = linprog(np.zeros(n), A_ub=Aub, b_ub=bub, bounds=[(None, None)] * n, method="highs")
    return r.status == 2


def farkas_certificate_exists(A, l, u):
    """min u'y+ - l'y-  s.t. A'(y+ - y-) = 0, sum(y) = 1, y >= 0 (y+ only on finite u, y- only on finite l) is negative."""
    m, n = A.shape
    fin_u, fin_l = u < INF, l > -INF
    c = np.concatenate([np.where(fin_u, u, 0.0), np.where(fin_l, -l, 0.0)])
    Aeq = np.vstack([np.hstack([A.T, -A.T]), np.ones((1, 2 * m))])
    beq = np.concatenate([np.zeros(n), [1.0]])
    bounds = [(0, None) if f else (0, 0) for f in fin_u] + [(0, None) if f else (0, 0) for f in fin_l]
    r = linprog(c, A_eq=Aeq, b_eq=beq, bounds=bounds, method="highs")
    return r.status == 0 and r.fun < -1e-9


def lp_unbounded_direction_exists(P, q, A, l, u):
    """min q'd  s.t. P d = 0, (A d)_i <= 0 where u_i finite, >= 0 where l_i finite, |d|_inf <= 1  is negative."""
    n = P.shape[0]
    fin_u, fin_l = u < INF, l > -INF
    Aub = np.vstack([A[fin_u], -A[fin_l]]) if (fin_u.any() or fin_l.any()) else None
    bub = np.zeros(Aub.shape[0]) if Aub is not None else None
    r = linprog(q, A_ub=Aub, b_ub=bub, A_eq=P, b_eq=np.zeros(n), bounds=[(-1, 1)] * n, method="highs")
    return r.status == 0 and r.fun < -1e-9


def solve_shared(P, A, l0, u0, q, l, u, kernel):
    s = sm.BatchedSolver(P, A, l0, u0, batch=q.shape[0], kernel=kernel, **EPS)
    s.update_gradient(q)
    if A.shape[0]:
        s.update_bounds(np.where(np.isfinite(l), l, np.sign(l) * np.inf), u)
    s.solve()
    x, y = s.solution()
    info = s.info()
    s.close()
    return x, y, info


def solve_instances(Ps, As, l0, u0, q, l, u):
    s = sm.BatchedSolver.batched(Ps, As, l0, u0, **EPS)
    s.update_gradient(q); s.update_bounds(l, u); s.solve()
    x, y = s.solution()
    info = s.info()
    s.close()
    return x, y, info


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(seed=hst.integers(0, 10_000), n=hst.integers(2, 16), mfac=hst.sampled_from([0.5, 1.0, 2.0]),
       kernel=hst.sampled_from([1, 2, 4, 5]))
def test_random_qps_are_kkt_optimal_shared_factor(seed, n, mfac, kernel):
    """hypothesis-driven random strictly convex QPs, every shared-factor kernel: each SOLVED instance is the exact optimum."""
    m = max(1, min(32, int(round(mfac * n))))
    P, q0, A, l0, u0 = random_qp(n, m, seed=seed)
    rng = np.random.default_rng(seed + 1)
    B = 8
    q = q0[None, :] + 0.3 * rng.standard_normal((B, n))
    sh = 0.15 * rng.standard_normal((B, m))
    l, u = l0[None, :] + sh, u0[None, :] + sh
    x, y, info = solve_shared(P, A, l0, u0, q, l, u, kernel)
    assert (info["status"] == sm.SOLVED).all()          # finite two-sided bounds with l < u, P > 0: always solvable
    checked = 0
    for b in range(B):
        Ax = A @ x[b]
        slack = np.minimum(u[b] - Ax, Ax - l[b])
        yact = np.abs(y[b]) > 1e-7 * max(np.abs(y[b]).max(), 1e-12)
        if (yact & (np.abs(y[b]) < 1e-3)).any() or ((~yact) & (slack < 1e-3)).any():
            continue                                     # (weakly active rows: the active set cannot be read off at eps = 1e-5)
        assert_kkt_optimal(P, q[b], A, l[b], u[b], x[b], y[b])
        checked += 1
    assume(checked > 0)


@settings(max_examples=6, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(seed=hst.integers(0, 10_000), n=hst.sampled_from([6, 17, 30]))
def test_random_qps_are_kkt_optimal_per_instance(seed, n):
    """the per-instance regime (own P_i, A_i per QP, batched Cholesky path): each SOLVED instance is the exact optimum."""
    m, B = 2 * n, 6
    _, _, _, l0, u0 = random_qp(n, m, seed=seed)
    data = [random_qp(n, m, seed=seed + 17 * (b + 1)) for b in range(B)]
    Ps, As, qs = (np.array([d[k] for d in data]) for k in (0, 2, 1))
    ls, us = np.array([d[3] for d in data]), np.array([d[4] for d in data])
    # the row classes are fixed by the setup bounds: all rows are inequalities here, as in the instance bounds
    x, y, info = solve_instances(Ps, As, np.full(m, -1.0), np.full(m, 1.0), qs, ls, us)
    assert (info["status"] == sm.SOLVED).all()
    checked = 0
    for b in range(B):
        Ax = As[b] @ x[b]
        slack = np.minimum(us[b] - Ax, Ax - ls[b])
        yact = np.abs(y[b]) > 1e-7 * max(np.abs(y[b]).max(), 1e-12)
        if (yact & (np.abs(y[b]) < 1e-3)).any() or ((~yact) & (slack < 1e-3)).any():
            continue
        assert_kkt_optimal(Ps[b], qs[b], As[b], ls[b], us[b], x[b], y[b])
        checked += 1
    assume(checked > 0)


def _infeasible_family(n, mp, B, seed):
    """[G; -G] rows (the reference's two-sided limit): G x <= u_top, -G x <= u_bot is empty iff u_top + u_bot < 0 on a row."""
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    P = M @ M.T / n + 0.5 * np.eye(n)
    G = rng.standard_normal((mp, n))
    A = np.vstack([G, -G])
    q = rng.standard_normal((B, n))
    u = 1.0 + rng.random((B, 2 * mp))
    bad = np.zeros(B, bool)
    bad[::3] = True
    rows = rng.integers(0, mp, B)
    for b in np.nonzero(bad)[0]:
        u[b, rows[b]] = -1.0 - rng.random(); u[b, mp + rows[b]] = -0.5 - rng.random()
    l = np.full((B, 2 * mp), -np.inf)
    return P, A, q, l, u, bad


@pytest.mark.parametrize("kernel", [1, 2, 4, 5])
@pytest.mark.parametrize("n,mp", [(3, 2), (10, 10), (16, 16)])
def test_primal_infeasible_status_is_proved_by_an_lp(kernel, n, mp):
    B = 24
    P, A, q, l, u, bad = _infeasible_family(n, mp, B, seed=n * 31 + mp)
    x, y, info = solve_shared(P, A, np.full(2 * mp, -np.inf), np.full(2 * mp, 2.0), q, l, u, kernel)
    assert np.array_equal(info["status"] == sm.PRIMAL_INFEASIBLE, bad) and (info["status"][~bad] == sm.SOLVED).all()
    lfin = np.full(2 * mp, -INF)
    for b in range(B):
        empty = lp_set_is_empty(A, lfin, u[b])
        assert empty == bool(bad[b])                       # the status is exactly what the LP proves
        if bad[b]:
            assert farkas_certificate_exists(A, lfin, u[b]) and np.isnan(x[b]).all() and np.isnan(y[b]).all()
        else:
            assert_kkt_optimal(P, q[b], A, lfin, u[b], x[b], y[b])


@pytest.mark.parametrize("n,mp", [(6, 5), (30, 30)])
def test_primal_infeasible_status_per_instance_regime(n, mp):
    B = 12
    Ps, As, qs, us, bads = [], [], [], [], []
    for b in range(B):
        P, A, q, l, u, bad = _infeasible_family(n, mp, 3, seed=1000 + 7 * b + n)
        k = b % 3                                            # instance 0 of each family is infeasible
        Ps.append(P); As.append(A); qs.append(q[k]); us.append(u[k]); bads.append(bad[k])
    Ps, As, qs, us, bads = map(np.array, (Ps, As, qs, us, bads))
    ls = np.full((B, 2 * mp), -np.inf)
    x, y, info = solve_instances(Ps, As, np.full(2 * mp, -np.inf), np.full(2 * mp, 2.0), qs, ls, us)
    assert np.array_equal(info["status"] == sm.PRIMAL_INFEASIBLE, bads) and (info["status"][~bads] == sm.SOLVED).all()
    lfin = np.full(2 * mp, -INF)
    for b in range(B):
        assert lp_set_is_empty(As[b], lfin, us[b]) == bool(bads[b])
        if not bads[b]:
            assert_kkt_optimal(Ps[b], qs[b], As[b], lfin, us[b], x[b], y[b])


@pytest.mark.parametrize("kernel", [1, 2, 4, 5])
def test_dual_infeasible_status_is_proved_by_an_lp(kernel):
    """P singular along e_n, the constraints leave that direction free upwards: q_n < 0 makes the QP unbounded."""
    n, m, B = 6, 4, 16
    rng = np.random.default_rng(3)
    M = rng.standard_normal((n - 1, n - 1))
    P = np.zeros((n, n)); P[: n - 1, : n - 1] = M @ M.T / n + 0.5 * np.eye(n - 1)
    A = rng.standard_normal((m, n)); A[:, n - 1] = np.abs(A[:, n - 1])      # moving up e_n only increases A x
    l0, u0 = np.full(m, -1.0), np.full(m, np.inf)                         # lower bounds only: up is free
    q = rng.standard_normal((B, n))
    q[:, n - 1] = np.where(np.arange(B) % 2 == 0, -1.0, 0.0)               # even instances unbounded, odd ones bounded (q_n = 0)
    l, u = np.tile(l0, (B, 1)), np.tile(u0, (B, 1))
    x, y, info = solve_shared(P, A, l0, u0, q, l, u, kernel)
    ufin = np.full(m, INF)
    for b in range(B):
        unb = lp_unbounded_direction_exists(P, q[b], A, l0, ufin)
        assert unb == (b % 2 == 0)
        assert (info["status"][b] == sm.DUAL_INFEASIBLE) == unb, (b, info["status"][b])
        if unb:
            assert np.isnan(x[b]).all()
        else:
            assert info["status"][b] == sm.SOLVED


def test_reference_config_file_byte_copy_parses_to_the_golden_matrices(repo_root, golden):
    """tests/golden/MPC_API.reference.json is a byte copy of the reference's config/MPC_API.json (dead keys A, B, C, D, Dd, Ts,
    t0 and the "____" separators included); smpc_mpc_create_from_json must accept it unchanged."""
    path = os.path.join(repo_root, "tests", "golden", "MPC_API.reference.json")
    raw = open(path).read()
    assert '"A"' in raw and "____" in raw          # it really is the unmodified file, not the repo's trimmed config
    mpc = sm.BatchedModelPredictiveControlAPI(path, batch=2, **EPS)
    assert (mpc.mpcWindow, mpc.N_S, mpc.n_variables, mpc.n_constraints) == (15, 4, 15, 30)
    for name in ("H", "Gbar", "Fx", "Fu", "Fr", "Sbar", "Ku", "W0", "Sx", "Su", "CAB"):
        g = np.array(golden[name], dtype=float).reshape(mpc.matrix(name).shape)
        assert np.abs(mpc.matrix(name) - g).max() <= 1e-12 * max(np.abs(g).max(), 1e-300), name
    c = golden["cases"][1]
    mpc.set_state(X=np.array([c["X"], c["X"]]), U=np.array([c["U"]] * 2), ref=np.array([c["xref"]] * 2))
    assert mpc.controllerStep()
    x, _ = mpc.solver.solution()
    assert np.abs(x[0] - np.array(c["x"])).max() <= 1e-7 * np.abs(c["x"]).max() and np.array_equal(x[0], x[1])
    mpc.close()
