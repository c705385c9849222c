"""GPU parity of the per-instance regime (every QP has its own P_i, A_i: batched Cholesky path, BASELINE config 4)."""
import os
import numpy as np
import pytest

import oracle
import solvempc_b200 as sm
from problems import c2_batch, c4_plants, random_qp
from test_gpu_parity import EPS, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,m,B", [(6, 9, 16), (30, 60, 48), (33, 40, 8)])
def test_random_per_instance_qps(n, m, B):
    Ps, As, qs, ls, us = [], [], [], [], []
    P0, q0, A0, l0, u0 = random_qp(n, m, seed=100)
    rng = np.random.default_rng(n)
    for b in range(B):
        P, q, A, _, _ = random_qp(n, m, seed=1000 + b)
        Ps.append(P); As.append(A); qs.append(q)
        sh = 0.1 * rng.standard_normal(m)
        ls.append(l0 + sh); us.append(u0 + sh)
    Ps, As, qs, ls, us = map(np.array, (Ps, As, qs, ls, us))
    ls[:, 0] = us[:, 0]                                   # an equality row everywhere
    l0 = l0.copy(); l0[0] = u0[0]
    s = sm.BatchedSolver.batched(Ps, As, l0, u0, **EPS)
    assert s.kernel_name == "admm_instance_kernel" and (s.n, s.m, s.batch) == (n, m, B)
    s.update_gradient(qs); s.update_bounds(ls, us); s.solve()
    x, y = s.solution(); info = s.info()
    xs, ys, st, it, ru = [], [], [], [], []
    for b in range(B):
        so = oracle.Solver(Ps[b], np.zeros(n), As[b], l0, u0, **EPS)
        if b == 0:
            D, E, c = so.scaling()
            Dd, Ed, cd = s.scaling()
            assert np.abs(Dd - D).max() <= 1e-15 * D.max() and np.abs(Ed - E).max() <= 1e-15 * E.max() and abs(cd - c) <= 1e-15 * c
        so.update_lin_cost(qs[b]); so.update_bounds(ls[b], us[b])
        r = so.solve()
        xs.append(r["x"]); ys.append(r["y"]); st.append(r["status"]); it.append(r["iter"]); ru.append(r["rho_updates"])
    assert np.array_equal(info["status"], np.array(st)) and np.array_equal(info["iter"], np.array(it))
    assert np.array_equal(info["rho_updates"], np.array(ru))
    assert rel_err(x, np.array(xs)) < 1e-6 and rel_err(y, np.array(ys)) < 1e-5
    # warm second solve (rho and iterates persist), as the oracle does per instance
    s.solve()
    assert (s.info()["iter"][np.array(st) == 1] == 25).all()
    s.close()


def test_c4_per_instance_plants_through_mpc_api(ref_mats):
    """Config 4: distinct linearised plants per controller, N = 30, assembly on device + batched Cholesky."""
    _, cfg = ref_mats
    N, B = 30, 96
    Ad, Bd = c4_plants(B, cfg, seed=2)
    conf = dict(Ad=Ad, Bd=Bd, Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=N, per_instance=1)
    mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=B, **EPS)
    assert (mpc.n_variables, mpc.n_constraints) == (30, 60) and mpc.solver.kernel_name == "admm_instance_pair_kernel"
    X, U, ref = c2_batch(B, seed=31)
    mpc.set_state(X=X, U=U, ref=ref)
    ok = mpc.controllerStep()
    x, _ = mpc.solver.solution(); info = mpc.solver.info()
    _, Uout = mpc.state()
    fd, ubd = mpc.step_vectors()
    nsolved = 0
    for b in range(B):
        mats = oracle.mpc_build(**{**cfg, "Ad": Ad[b], "Bd": Bd[b], "N": N})
        if b % 16 == 0:
            for name in ("H", "Gbar", "Fx", "Fu", "Fr"):
                assert np.abs(mpc.matrix(name, b) - mats[name]).max() <= 1e-11 * np.abs(mats[name]).max(), name
        f, ub = oracle.mpc_step_vectors(mats, X[b], U[b], ref[b])
        assert np.abs(fd[b] - f).max() <= 1e-11 * np.abs(f).max() and np.abs(ubd[b] - ub).max() <= 1e-12 * np.abs(ub).max()
        so = oracle.Solver(mats["H"], np.zeros(N), mats["Gbar"], mats["lb"], mats["W0"], **EPS)
        so.update_lin_cost(f); so.update_upper_bound(ub)
        r = so.solve()
        assert info["status"][b] == r["status"] and info["iter"][b] == r["iter"], b
        assert rel_err(x[b], r["x"]) < 1e-6
        if r["status"] == 1:
            nsolved += 1
            assert abs(Uout[b] - (U[b] + r["x"][0])) < 1e-8
        else:
            assert Uout[b] == U[b]
    assert ok == (nsolved == B) and nsolved >= B // 2
    mpc.close()


@pytest.mark.parametrize("paired", [False, True])
def test_bounds_that_change_a_row_class_fall_back_to_the_full_factorisation(ref_mats, paired):
    """The factorisation prepared at create time (M = S0 + rho T for the setup bounds' row classes, M(rho0)^-1) must not be
    used when an instance's own bounds turn an inequality row into an equality row: rho_vec changes, OSQP refactors."""
    _, cfg = ref_mats
    B = 24
    if paired:      # [G; -G] rows (the reference's constraint form): per-instance plants through the MPC builders
        N = 12
        Ad, Bd = c4_plants(B, cfg, seed=5)
        mats = [oracle.mpc_build(**{**cfg, "Ad": Ad[b], "Bd": Bd[b], "N": N}) for b in range(B)]
        Ps, As = np.array([m["H"] for m in mats]), np.array([m["Gbar"] for m in mats])
        n, m = N, 2 * N
        l0, u0 = mats[0]["lb"], mats[0]["W0"]
        X, U, ref = c2_batch(B, seed=77)
        qs, us = [], []
        for b in range(B):
            f, ub = oracle.mpc_step_vectors(mats[b], X[b], U[b], ref[b])
            qs.append(f); us.append(ub)
        qs, us = np.array(qs), np.array(us)
        ls = np.tile(l0, (B, 1))
        ls[:, 3] = us[:, 3] = 0.0              # row 3 becomes an equality (its pair row 3 + N stays an inequality)
    else:
        n, m = 10, 16
        Ps, As, qs = [], [], []
        l0, u0 = np.full(m, -2.0), np.full(m, 2.0)   # wide enough to be feasible for every instance
        for b in range(B):
            P, q, A, _, _ = random_qp(n, m, seed=3000 + b)
            Ps.append(P); As.append(A); qs.append(q)
        Ps, As, qs = map(np.array, (Ps, As, qs))
        ls, us = np.tile(l0, (B, 1)), np.tile(u0, (B, 1))
        ls[:, 2] = us[:, 2] = 0.25             # inequality -> equality
        ls[:, 5], us[:, 5] = -np.inf, np.inf   # inequality -> free
    s = sm.BatchedSolver.batched(Ps, As, l0, u0, **EPS)
    assert s.row_pairs == (m // 2 if paired else 0)
    s.update_gradient(qs); s.update_bounds(ls, us); s.solve()
    x, y = s.solution(); info = s.info()
    for b in range(B):
        so = oracle.Solver(Ps[b], np.zeros(n), As[b], l0, u0, **EPS)
        so.update_lin_cost(qs[b]); so.update_bounds(ls[b], us[b])
        r = so.solve()
        assert info["status"][b] == r["status"] and info["iter"][b] == r["iter"], b
        assert rel_err(x[b], r["x"]) < 1e-6
    # back to the setup classes on the same handle: the prepared factorisation applies again, same answers as a fresh handle
    s.update_bounds(np.tile(l0, (B, 1)), np.tile(u0, (B, 1)) if not paired else us * 0 + u0)
    s.set_cold_solves(True); s.solve()
    x2, _ = s.solution()
    s2 = sm.BatchedSolver.batched(Ps, As, l0, u0, **EPS)
    s2.update_gradient(qs); s2.solve()
    x3, _ = s2.solution()
    assert np.array_equal(x2, x3) and np.isfinite(x3).all()
    assert (info["status"] == 1).mean() >= 0.5     # (forcing an equality makes a few of the paired instances infeasible: status parity above)
    s.close(); s2.close()


@pytest.mark.parametrize("N", [12, 15, 30, 32])
def test_pair_kernel_matches_the_one_warp_kernel_and_the_oracle(ref_mats, monkeypatch, N):
    """[G; -G] instances run on the two-warp TMA-staged kernel (admm_instance_pair.cu); SMPC_INSTANCE_NO_PAIR_KERNEL keeps them on
    the one-warp register kernel.  Same statuses and iteration counts, solutions to round-off; warm second solve (the stored
    factorisation of the final rho is reused) and a cold third solve (rho0: the prepared pack again) agree as well."""
    _, cfg = ref_mats
    B = 40
    Ad, Bd = c4_plants(B, cfg, seed=9)
    conf = dict(Ad=Ad, Bd=Bd, Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=N, per_instance=1)
    X, U, ref = c2_batch(B, seed=13)
    out = {}
    for name, env in (("pair", None), ("warp", "1")):
        if env:
            monkeypatch.setenv("SMPC_INSTANCE_NO_PAIR_KERNEL", env)
        else:
            monkeypatch.delenv("SMPC_INSTANCE_NO_PAIR_KERNEL", raising=False)
        mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=B, **EPS)
        assert mpc.solver.kernel_name == ("admm_instance_pair_kernel" if name == "pair" else "admm_instance_kernel")
        res = []
        mpc.set_state(X=X, U=U, ref=ref)
        mpc.controller_step_async()
        res.append((mpc.solver.solution(), mpc.solver.info(), mpc.state()[1]))
        mpc.set_state(X=X * 0.9)
        mpc.controller_step_async()                      # warm: iterates, rho and the factorisation persist
        res.append((mpc.solver.solution(), mpc.solver.info(), mpc.state()[1]))
        mpc.solver.set_cold_solves(True)
        mpc.set_state(X=X, U=U, ref=ref)
        mpc.controller_step_async()                      # cold again: bitwise the first solve
        res.append((mpc.solver.solution(), mpc.solver.info(), mpc.state()[1]))
        out[name] = res
        mpc.close()
    for k in range(3):
        (xp, yp), ip, up = out["pair"][k]
        (xw, yw), iw, uw = out["warp"][k]
        assert np.array_equal(ip["status"], iw["status"]) and np.array_equal(ip["iter"], iw["iter"]), k
        assert np.array_equal(ip["rho_updates"], iw["rho_updates"])
        assert rel_err(xp, xw) < 1e-7 and rel_err(yp, yw) < 1e-6
        assert np.abs(up - uw).max() < 1e-9
    assert np.array_equal(out["pair"][0][0][0], out["pair"][2][0][0])
    # and the oracle on a few instances
    (xp, _), ip, _ = out["pair"][0]
    for b in range(0, B, 8):
        mats = oracle.mpc_build(**{**cfg, "Ad": Ad[b], "Bd": Bd[b], "N": N})
        f, ub = oracle.mpc_step_vectors(mats, X[b], U[b], ref[b])
        so = oracle.Solver(mats["H"], np.zeros(N), mats["Gbar"], mats["lb"], mats["W0"], **EPS)
        so.update_lin_cost(f); so.update_upper_bound(ub)
        r = so.solve()
        assert ip["status"][b] == r["status"] and ip["iter"][b] == r["iter"]
        assert rel_err(xp[b], r["x"]) < 1e-6


def test_pair_kernel_against_the_oracles_per_plant_controllers(ref_mats):
    """4 096 distinct plants (N = 30) through the two-warp kernel -- rho updates from the per-instance pencil
    eigen-decomposition -- against one oracle controller per plant (oracle/batch_drivers.c): every status and every iteration
    count identical, SOLVED solutions equal to round-off."""
    _, cfg = ref_mats
    B, N = 4096, 30
    Ad, Bd = c4_plants(B, cfg, seed=5)
    X, U, ref = c2_batch(B, seed=105)
    conf = dict(Ad=Ad, Bd=Bd, Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=N, per_instance=1)
    mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=B, **EPS)
    assert mpc.solver.kernel_name == "admm_instance_pair_kernel"
    mpc.set_state(X=X, U=U, ref=ref)
    mpc.controller_step_async()
    x, _ = mpc.solver.solution(); info = mpc.solver.info()
    mpc.close()
    out = oracle.plant_batch(cfg, Ad, Bd, X, U, ref, N, settings=oracle.default_settings(**EPS), nthreads=os.cpu_count() or 1)
    assert np.array_equal(info["status"], out["status"]) and np.array_equal(info["iter"], out["iter"])
    ok = out["status"] == 1
    assert ok.mean() > 0.95 and info["rho_updates"].mean() > 0.5          # (the rho-update path is exercised)
    assert (np.abs(x[ok] - out["x"][ok]).max(axis=1) / np.maximum(np.abs(out["x"][ok]).max(axis=1), 1e-9)).max() < 1e-8
