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
    assert mpc.solver.kernel_name == "admm_instance_kernel" and mpc.solver.row_pairs == N
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
