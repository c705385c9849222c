"""CPU checks of the boundary: the C-ABI library loads, exports every symbol include/solvempc_b200.h
declares, fails loudly without a device (no CPU fallback), and the host-side plan maths is right."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
import solvempc_b200 as sm
from solvempc_b200 import _lib as L
from problems import c2_batch, random_qp


def _header_symbols(root):
    txt = open(os.path.join(root, "include", "solvempc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(smpc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(repo_root):
    syms = _header_symbols(repo_root)
    assert len(syms) >= 40
    lib = C.CDLL(sm.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(L.SIGNATURES) == syms, set(syms) ^ set(L.SIGNATURES)


def test_version_and_defaults():
    assert b"sm_100a" in sm.lib().smpc_version()
    s = sm.default_settings()
    assert (s.rho, s.sigma, s.alpha, s.eps_abs, s.max_iter, s.check_termination, s.scaling) == (0.1, 1e-6, 1.6, 1e-3, 4000, 25, 10)
    assert s.adaptive_rho == 1 and s.adaptive_rho_interval == 25 and s.warm_start == 1


def test_argument_errors_are_reported_without_a_device():
    with pytest.raises(sm.SolveMpcError) as e:
        sm.BatchedSolver(np.eye(2), np.eye(2), np.ones(2), np.zeros(2), batch=1)   # l > u
    assert e.value.code == L.ERR_DATA
    with pytest.raises(sm.SolveMpcError) as e:
        sm.BatchedSolver(-np.eye(2), np.eye(2), batch=1)                            # not PD
    assert e.value.code == L.ERR_DATA
    with pytest.raises(sm.SolveMpcError) as e:
        sm.BatchedSolver(np.eye(2), np.eye(2), batch=1, rho=-1.0)
    assert e.value.code == L.ERR_ARG
    with pytest.raises(sm.SolveMpcError) as e:
        sm.BatchedModelPredictiveControlAPI("/nonexistent/MPC_API.json")
    assert e.value.code == L.ERR_IO


def test_malformed_config_is_an_io_error(tmp_path):
    bad = tmp_path / "MPC_API.json"
    bad.write_text('{"Ad": [[1,2],[3]], "Bd": [[1],[2]]}')
    with pytest.raises(sm.SolveMpcError) as e:
        sm.BatchedModelPredictiveControlAPI(str(bad))
    assert e.value.code == L.ERR_IO
    bad.write_text('{"Ad": [[1,0],[0,1]], "Bd": [[1],[2]], "Cd": [[1,0]], "Dd": [[0]], "K": [[1,2,3]], "Q": [[1]], "R": [[1]], "RD": [1], "xref": 0}')
    with pytest.raises(sm.SolveMpcError) as e:
        sm.BatchedModelPredictiveControlAPI(str(bad))
    assert e.value.code == L.ERR_IO and "K" in str(e.value)


@pytest.mark.skipif(sm.lib().smpc_device_count() > 0, reason="a CUDA device is present")
def test_no_cpu_fallback(ref_mats):
    m, _ = ref_mats
    with pytest.raises(sm.SolveMpcError) as e:
        sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=4)
    assert e.value.code == L.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_plan_scaling_matches_oracle_bitwise(ref_mats):
    m, _ = ref_mats
    pl = sm.shared_plan_inspect(m["H"], m["Gbar"], m["lb"], m["W0"])
    s = oracle.Solver(m["H"], np.zeros(15), m["Gbar"], m["lb"], m["W0"])
    D, E, c = s.scaling()
    assert np.array_equal(pl["D"], D) and np.array_equal(pl["E"], E) and pl["c"] == c
    assert (pl["ctype"] == 0).all()


def _plan_identities(P, A, l0, u0, tol=1.0, **kw):
    s = oracle.Solver(P, np.zeros(P.shape[0]), A, l0, u0, **kw)
    Pb, Ab = s.scaled_data()
    pl = sm.shared_plan_inspect(P, A, l0, u0, **kw)
    n = P.shape[0]
    sigma = 1e-6
    V, lam = pl["V"], pl["lam"]
    kap = np.where(pl["ctype"] == 0, 1.0, np.where(pl["ctype"] == 1, 1e3, 0.0))
    free = pl["ctype"] == -1
    S = Pb + sigma * np.eye(n) + 1e-6 * Ab[free].T @ Ab[free]
    T = Ab.T @ (kap[:, None] * Ab)
    assert np.abs(V.T @ S @ V - np.eye(n)).max() < 1e-12 * tol
    assert np.abs(V.T @ T @ V - np.diag(lam)).max() < 1e-10 * max(1.0, lam.max())
    assert np.abs(pl["SG"] - sigma * V.T @ V).max() < 1e-18 * tol * tol
    assert np.abs(pl["W"] - Ab @ V).max() < 1e-12 * tol * max(1.0, np.abs(Ab).max())
    assert np.abs(pl["PVT"].T - Pb @ V).max() < 1e-12 * tol
    assert np.abs(pl["VinvT"].T @ V - np.eye(n)).max() < 1e-11 * tol
    for rho in (1e-6, 0.1, 37.0, 1e6):   # M(rho)^-1 = V diag(1/(1+rho lam)) V'
        M = S + rho * T
        Minv = V @ np.diag(1.0 / (1.0 + rho * lam)) @ V.T
        assert np.abs(M @ Minv - np.eye(n)).max() < 1e-9 * tol
    return pl


def test_plan_pencil_identities(ref_mats):
    m, _ = ref_mats
    _plan_identities(m["H"], m["Gbar"], m["lb"], m["W0"])
    for seed, (n, mm) in enumerate([(5, 3), (12, 20), (40, 70)]):
        P, q, A, l, u = random_qp(n, mm, seed)
        l[: mm // 4] = u[: mm // 4]            # equality rows
        l[mm // 4] = -np.inf; u[mm // 4] = np.inf   # a free row
        pl = _plan_identities(P, A, l, u)
        assert (pl["ctype"][: mm // 4] == 1).all() and pl["ctype"][mm // 4] == -1
    _plan_identities(*[random_qp(8, 10, 5)[k] for k in (0, 2, 3, 4)], scaling=0)


def test_plan_pencil_identities_at_config3_size():
    """n = 200, m = 400 (BASELINE config 3): the Jacobi sweep must converge (not run out its sweep limit with a residual) and
    build in seconds, not minutes."""
    import time
    m3 = oracle.mimo_build(**oracle.load_mimo_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "quadrotor.json")))
    t0 = time.time()
    pl = _plan_identities(m3["H"], m3["A"], m3["lb"], m3["ub"], tol=2e3)   # cond(S) = 5e7 here: residuals ~ cond * eps
    assert time.time() - t0 < 60.0
    assert pl["lam"].min() > 0


def _emulate_device_iteration(pl, Ab, qbar, lbar, ubar, eps, c, D, E, rho0=0.1, sigma=1e-6, alpha=1.6, max_iter=4000):
    """numpy statement of the kernels' iteration (admm_shared_*.cu) on the plan operators."""
    V, lam, W, SG = pl["V"], pl["lam"], pl["W"], pl["SG"]
    PV = pl["PVT"].T
    ct = pl["ctype"]
    n, m = V.shape[0], W.shape[0]
    xi, z, y, rho = np.zeros(n), np.zeros(m), np.zeros(m), rho0
    qh = V.T @ qbar
    rvec = lambda r: np.where(ct == 0, r, np.where(ct == 1, 1e3 * r, 1e-6))
    Dinv, Einv, cinv = 1 / D, 1 / E, 1 / c
    for it in range(1, max_iter + 1):
        rv = rvec(rho)
        w = rv * z - y
        t = (SG @ xi + W.T @ w - qh) / (1 + rho * lam)
        zt = W @ t
        xi = alpha * t + (1 - alpha) * xi
        zr = alpha * zt + (1 - alpha) * z
        zn = np.minimum(np.maximum(zr + y / rv, lbar), ubar)
        y = y + rv * (zr - zn)
        z = zn
        if it % 25 == 0:
            x, Ax, Px, Aty = V @ xi, W @ xi, PV @ xi, Ab.T @ y
            rp, rd = Ax - z, qbar + Px + Aty
            pri, dua = np.abs(Einv * rp).max(), cinv * np.abs(Dinv * rd).max()
            ep = eps + eps * max(np.abs(Einv * z).max(), np.abs(Einv * Ax).max())
            ed = eps + eps * cinv * max(np.abs(Dinv * qbar).max(), np.abs(Dinv * Aty).max(), np.abs(Dinv * Px).max())
            if pri < ep and dua < ed:
                return D * x, E * y / c, 1, it
            pn = np.abs(rp).max() / (max(np.abs(z).max(), np.abs(Ax).max()) + 1e-10)
            dn = np.abs(rd).max() / (max(np.abs(qbar).max(), np.abs(Aty).max(), np.abs(Px).max()) + 1e-10)
            rn = min(max(rho * np.sqrt(pn / (dn + 1e-10)), 1e-6), 1e6)
            if rn > 5 * rho or rn < rho / 5:
                rho = rn
    return D * (V @ xi), E * y / c, -2, max_iter


def test_plan_iteration_follows_the_oracle_path(ref_mats):
    """The plan-coordinate iteration the kernels run is OSQP's iteration: same iterates, same iteration
    counts, for per-instance adapted rho, with ONE shared set of operators."""
    m, _ = ref_mats
    X, U, ref = c2_batch(48, seed=3)
    f, ub = oracle.mpc_batch_vectors(m, X, U, ref)
    out = oracle.solve_batch(m["H"], m["Gbar"], m["lb"], m["W0"], f, ub, nthreads=2, eps_abs=1e-5, eps_rel=1e-5)
    s = oracle.Solver(m["H"], np.zeros(15), m["Gbar"], m["lb"], m["W0"])
    _, Ab = s.scaled_data()
    pl = sm.shared_plan_inspect(m["H"], m["Gbar"], m["lb"], m["W0"])
    D, E, c = pl["D"], pl["E"], pl["c"]
    for b in range(48):
        x, y, st, it = _emulate_device_iteration(pl, Ab, c * D * f[b], np.full(30, -np.inf), E * ub[b], 1e-5, c, D, E)
        assert st == out["status"][b] and it == out["iter"][b]
        assert np.abs(x - out["x"][b]).max() <= 1e-10 * max(np.abs(out["x"][b]).max(), 1e-12)
        assert np.abs(y - out["y"][b]).max() <= 1e-9 * max(np.abs(out["y"][b]).max(), 1e-9)
