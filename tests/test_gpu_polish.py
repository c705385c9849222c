"""GPU parity of solution polishing (OSQP polish.c; off in the reference, cpp:51-52).

The polished point is pinned two ways: (i) against the exact KKT solve of the known active set (solver independent, cases
A-D of SURVEY 8c and random QPs whose active set the polished oracle and numpy agree on), (ii) against the oracle's polish
(status_polish identical, iterates to round-off).  Tolerances: 1e-9 relative on x and y (two dense LDL' factorisations of a
quasi-definite matrix in different elimination orders + 3 refinement steps)."""
import numpy as np
import pytest

import oracle
import solvempc_b200 as sm
from problems import KNOWN_CASES, c2_batch, c4_plants, random_qp
from test_gpu_parity import EPS, rel_err

pytestmark = pytest.mark.gpu


def _oracle_polished(P, q0, A, l0, u0, q, l, u, **kw):
    so = oracle.Solver(P, q0, A, l0, u0, polish=1, **{**EPS, **kw})
    so.update_lin_cost(q)
    if l is not None:
        so.update_bounds(l, u)
    else:
        so.update_upper_bound(u)
    return so.solve()


@pytest.mark.parametrize("kernel", [1, 2, 4, 5])
def test_known_answers_polished_to_machine_precision(ref_mats, kernel):
    """Cases A-D through the MPC layer with polish on: dU* and y equal the exact KKT answer to ~1e-12, far inside eps."""
    m, cfg = ref_mats
    mpc = sm.BatchedModelPredictiveControlAPI(cfg, batch=len(KNOWN_CASES), kernel=kernel, **EPS)
    mpc.solver.set_polish(True)
    X = np.array([c["X"] for c in KNOWN_CASES]); U = np.array([c["U"] for c in KNOWN_CASES]); ref = np.array([c["ref"] for c in KNOWN_CASES])
    mpc.set_state(X=X, U=U, ref=ref)
    mpc.controllerStep()
    x, y = mpc.solver.solution(); info = mpc.solver.info(); pol = mpc.solver.polish_status()
    _, Uout = mpc.state()
    assert (info["status"] == 1).all() and (pol == 1).all()
    for b, c in enumerate(KNOWN_CASES):
        f, ub = oracle.mpc_step_vectors(m, c["X"], c["U"], c["ref"])
        xe, ye = oracle.exact_qp_active_set(m["H"], f, m["Gbar"], ub, c["active"])
        assert np.abs(x[b] - xe).max() <= 1e-12 * max(1.0, np.abs(xe).max()), (c["name"], np.abs(x[b] - xe).max())
        assert np.abs(y[b] - ye).max() <= 1e-11 * max(1.0, np.abs(ye).max()), (c["name"], np.abs(y[b] - ye).max())
        assert abs(x[b, 0] - c["du0"]) <= 1e-12 and abs(info["obj"][b] - c["obj"]) <= 1e-12 * max(1.0, abs(c["obj"]))
        assert abs(Uout[b] - (c["U"] + xe[0])) <= 1e-12          # U += dU[0] uses the POLISHED first move (cpp:105)
        r = _oracle_polished(m["H"], np.zeros(15), m["Gbar"], m["lb"], m["W0"], f, None, ub)
        assert r["status_polish"] == 1 and info["iter"][b] == r["iter"]
        assert info["pri_res"][b] <= 1e-12 and info["dua_res"][b] <= 1e-12
    mpc.close()


@pytest.mark.parametrize("n,m,kernel", [(8, 12, 2), (16, 32, 2), (20, 30, 1), (40, 64, 4), (100, 160, 4)])
def test_random_qps_polish_matches_oracle(n, m, kernel):
    """Shared-factor regime, two-sided bounds with an equality row; n + m = 260 takes the global-scratch path."""
    B = 24
    P, q0, A, l0, u0 = random_qp(n, m, seed=7)
    l0 = l0.copy(); l0[0] = u0[0]
    rng = np.random.default_rng(n * 100 + m)
    q = q0 + 0.5 * rng.standard_normal((B, n))
    sh = 0.1 * rng.standard_normal((B, m))
    l, u = l0 + sh, u0 + sh
    s = sm.BatchedSolver(P, A, l0, u0, batch=B, q0=q0, kernel=kernel, **EPS)
    s.set_polish(True)
    s.update_gradient(q); s.update_bounds(l, u); s.solve()
    x, y = s.solution(); info = s.info(); pol = s.polish_status()
    for b in range(B):
        r = _oracle_polished(P, q0, A, l0, u0, q[b], l[b], u[b])
        assert info["status"][b] == r["status"] and info["iter"][b] == r["iter"] and pol[b] == r["status_polish"], b
        assert rel_err(x[b], r["x"]) < 1e-9 and rel_err(y[b], r["y"]) < 1e-9, (b, rel_err(x[b], r["x"]), rel_err(y[b], r["y"]))
        if pol[b] == 1:
            k = oracle.kkt_report(P, q[b], A, l[b], u[b], x[b], y[b])
            assert k["stationarity"] < 1e-10 and k["infeasibility"] < 1e-10 and k["complementarity"] < 1e-10, (b, k)
            assert abs(info["pri_res"][b] - r["pri_res"]) <= 1e-12 and abs(info["dua_res"][b] - r["dua_res"]) <= 1e-10
    assert (pol == 1).sum() >= B // 2
    # the polished (x, z, y) is the next solve's warm start, as in OSQP: the second solve converges at its first check
    s.solve()
    assert (s.info()["iter"][pol == 1] == 25).all()
    s.close()


def test_polish_off_by_default_and_status_zero_when_not_solved():
    P, q0, A, l0, u0 = random_qp(10, 14, seed=3)
    A2 = A.copy(); A2[3] = A2[2]
    l0, u0 = l0.copy(), u0.copy(); l0[3], u0[3] = l0[2], u0[2]
    l = np.tile(l0, (4, 1)); u = np.tile(u0, (4, 1))
    l[1, 2], u[1, 2], l[1, 3], u[1, 3] = 5.0, 6.0, -6.0, -5.0  # A2[2] x in [5, 6] and in [-6, -5]: primal infeasible
    s = sm.BatchedSolver(P, A2, l0, u0, batch=4, q0=q0, **EPS)
    s.update_bounds(l, u); s.solve()
    assert (s.polish_status() == 0).all()
    s.set_polish(True)
    s.solve()
    st, pol = s.info()["status"], s.polish_status()
    assert st[1] == sm.PRIMAL_INFEASIBLE and pol[1] == 0 and (pol[[0, 2, 3]] == 1).all()
    assert np.isnan(s.solution()[0][1]).all()
    s.set_polish(False)
    s.solve()
    s.close()


def test_polish_per_instance_regime(ref_mats):
    """Config-4 style: distinct plants per controller, the pair kernel's x-space state, per-instance scaling."""
    _, cfg = ref_mats
    N, B = 30, 32
    Ad, Bd = c4_plants(B, cfg, seed=5)
    conf = dict(Ad=Ad, Bd=Bd, Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=N, per_instance=1)
    mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=B, **EPS)
    mpc.solver.set_polish(True)
    X, U, ref = c2_batch(B, seed=77)
    mpc.set_state(X=X, U=U, ref=ref)
    mpc.controllerStep()
    x, y = mpc.solver.solution(); info = mpc.solver.info(); pol = mpc.solver.polish_status()
    _, Uout = mpc.state()
    for b in range(B):
        mats = oracle.mpc_build(**{**cfg, "Ad": Ad[b], "Bd": Bd[b], "N": N})
        f, ub = oracle.mpc_step_vectors(mats, X[b], U[b], ref[b])
        r = _oracle_polished(mats["H"], np.zeros(N), mats["Gbar"], mats["lb"], mats["W0"], f, None, ub)
        assert info["status"][b] == r["status"] and info["iter"][b] == r["iter"] and pol[b] == r["status_polish"], b
        assert rel_err(x[b], r["x"]) < 1e-9 and rel_err(y[b], r["y"]) < 1e-8, (b, rel_err(x[b], r["x"]), rel_err(y[b], r["y"]))
        if r["status"] == 1:
            assert abs(Uout[b] - (U[b] + r["x"][0])) < 1e-10
    assert (pol == 1).sum() >= B // 2
    mpc.close()
