"""BASELINE-size runs of configs 4 and 5 on the GPU, checked through size-independent properties and a sample of
instances against the CPU oracle (configs 2 and 3 at full size: test_gpu_parity.py / test_gpu_mimo.py)."""
import os

import numpy as np
import pytest

import oracle
import solvempc_b200 as sm
from problems import c2_batch, c4_plants

pytestmark = pytest.mark.gpu

EPS = dict(eps_abs=1e-5, eps_rel=1e-5)


def test_c5_closed_loop_at_full_size(ref_mats):
    """65 536 warm-started controllers, N = 100: every solve of every step ends SOLVED in 25-iteration multiples, the state
    stays bounded, and a sample of controllers follows the oracle's closed loop step by step."""
    _, cfg = ref_mats
    N, B, steps, amp, period = 100, 65536, 4, 0.1, 200
    conf = dict(Ad=cfg["Ad"], Bd=cfg["Bd"], Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=N)
    X0, U0, _ = c2_batch(B, seed=31)
    X0, U0 = X0 * 0.2, U0 * 0.1
    phase = np.random.default_rng(5).integers(0, period, B).astype(np.int32)
    mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=B, **EPS)
    assert mpc.solver.kernel_name == "admm_shared_tile_kernel" and mpc.solver.row_pairs == N
    mpc.set_state(X=X0, U=U0, ref=np.zeros(B))
    bad, iters = mpc.closed_loop(steps, amp, period, phase)
    assert bad == 0 and iters % 25 == 0 and iters >= 25 * B * steps
    X, U = mpc.state()
    assert np.isfinite(X).all() and np.abs(U).max() < 255.0 + 1e-6
    # oracle closed loop for a sample (same warm-start semantics: iterates and rho persist per controller)
    mats = oracle.mpc_build(**{**cfg, "N": N})
    for b in range(0, B, B // 8):
        so = oracle.Solver(mats["H"], np.zeros(N), mats["Gbar"], mats["lb"], mats["W0"], **EPS)
        x, u = X0[b].copy(), float(U0[b])
        for k in range(steps):
            r = amp if 2 * ((k + int(phase[b])) % period) < period else -amp
            f, ub = oracle.mpc_step_vectors(mats, x, u, r)
            so.update_lin_cost(f); so.update_upper_bound(ub)
            res = so.solve()
            assert res["status"] == 1
            u += res["x"][0]
            x = cfg["Ad"] @ x + cfg["Bd"] * u
        assert np.abs(X[b] - x).max() <= 1e-6 * max(1.0, np.abs(x).max()) and abs(U[b] - u) <= 1e-6 * max(1.0, abs(u))
    mpc.close()


def test_c4_per_instance_plants_at_full_size(ref_mats):
    """65 536 distinct plants, N = 30: statuses, multiples of 25, bounded controls, and a sample against the oracle."""
    _, cfg = ref_mats
    N, B = 30, 65536
    Ad, Bd = c4_plants(B, cfg, seed=2)
    conf = dict(Ad=Ad, Bd=Bd, Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=N, per_instance=1)
    mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=B, **EPS)
    assert mpc.solver.kernel_name == "admm_instance_pair_kernel" and mpc.solver.row_pairs == N
    X, U, ref = c2_batch(B, seed=31)
    mpc.set_state(X=X, U=U, ref=ref)
    mpc.controller_step_async()
    info = mpc.solver.info()
    x, _ = mpc.solver.solution()
    solved = info["status"] == 1
    assert solved.mean() > 0.95 and set(np.unique(info["status"])) <= {1, -3, -2, 2, 3}
    assert (info["iter"][solved] % 25 == 0).all()
    _, Uo = mpc.state()
    assert np.array_equal(Uo[solved], (U + x[:, 0])[solved]) and np.array_equal(Uo[~solved], U[~solved])
    for b in range(0, B, B // 16):
        mats = oracle.mpc_build(**{**cfg, "Ad": Ad[b], "Bd": Bd[b], "N": N})
        f, ub = oracle.mpc_step_vectors(mats, X[b], U[b], ref[b])
        so = oracle.Solver(mats["H"], np.zeros(N), mats["Gbar"], mats["lb"], mats["W0"], **EPS)
        so.update_lin_cost(f); so.update_upper_bound(ub)
        r = so.solve()
        assert info["status"][b] == r["status"] and info["iter"][b] == r["iter"]
        if r["status"] == 1:
            assert np.abs(x[b] - r["x"]).max() <= 1e-6 * max(1e-9, np.abs(r["x"]).max())
    mpc.close()


def test_c3_quadrotor_at_full_size(repo_root):
    """One GPU's share of the 1M-instance batch of config 3 (131 072 quadrotor QPs, n = 200, m = 400): solver-independent
    KKT properties of every instance; with a sample against the oracle."""
    from problems import c3_batch
    path = os.path.join(repo_root, "config", "quadrotor.json")
    B = 131072
    x0, xr = c3_batch(B, seed=11)
    mpc = sm.BatchedMimoMPC(path, batch=B, **EPS)
    assert mpc.solver.row_pairs == 200
    mpc.set_state(x0=x0, xr=xr)
    assert mpc.controllerStep()
    x, y = mpc.solver.solution()
    info = mpc.solver.info()
    assert (info["iter"] % 25 == 0).all() and info["iter"].max() <= 500
    H, A, ub = mpc.matrix("H"), mpc.matrix("A"), mpc.matrix("ub")
    q = oracle.mimo_batch_vectors(dict(Fx=mpc.matrix("Fx"), Fr=mpc.matrix("Fr")), x0, xr)
    Px, Aty, Ax = x @ H, y @ A, x @ A.T
    stat = np.abs(Px + q + Aty).max(axis=1)
    assert (stat <= 1e-5 + 1e-5 * np.maximum(np.maximum(np.abs(Px).max(axis=1), np.abs(Aty).max(axis=1)), np.abs(q).max(axis=1))).all()
    assert ((Ax - ub).max(axis=1) <= 1e-5 + 1e-5 * np.abs(Ax).max(axis=1)).all()
    assert (y >= -1e-9).all()
    m = oracle.mimo_build(**oracle.load_mimo_config(path))
    idx = np.arange(0, B, B // 8)
    ora = oracle.solve_batch(H, A, m["lb"], ub, q[idx], np.tile(ub, (len(idx), 1)), nthreads=os.cpu_count() or 1, **EPS)
    assert np.array_equal(info["status"][idx], ora["status"]) and np.array_equal(info["iter"][idx], ora["iter"])
    assert (np.abs(x[idx] - ora["x"]).max(axis=1) / np.abs(ora["x"]).max(axis=1)).max() < 1e-4
    mpc.close()
