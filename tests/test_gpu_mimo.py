"""GPU parity of the multi-input MPC layer (BASELINE config 3: 12-state / 4-input quadrotor, N = 50 -> n = 200,
m = 400) through the C ABI: device assembly against the numpy oracle, the batched solve against the CPU oracle
on the same seeded inputs, and -- at a larger batch -- solver-independent KKT properties.

Tolerance (north star): u0 and the whole trajectory within 1e-4 relative, same status, same active set.
"""
import os

import numpy as np
import pytest

import oracle
import solvempc_b200 as sm
from problems import c3_batch

pytestmark = pytest.mark.gpu

EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
REL = 1e-4


@pytest.fixture(scope="module")
def quad(repo_root):
    path = os.path.join(repo_root, "config", "quadrotor.json")
    cfg = oracle.load_mimo_config(path)
    return path, cfg, oracle.mimo_build(**cfg)


def test_device_assembly_matches_numpy_oracle(quad):
    path, cfg, m = quad
    mpc = sm.BatchedMimoMPC(path, batch=2, **EPS)
    assert (mpc.horizon, mpc.nx, mpc.nu, mpc.n_variables, mpc.n_constraints) == (50, 12, 4, 200, 400)
    for name in ("H", "Fx", "Fr", "Su", "Sx", "A", "ub"):
        got, want = mpc.matrix(name), m[name]
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), name
    assert np.array_equal(mpc.matrix("H"), mpc.matrix("H").T)
    # the dict form of the config gives the same QP
    mpc2 = sm.BatchedMimoMPC(dict(cfg, horizon=cfg["N"]), batch=2, **EPS)
    assert np.array_equal(mpc2.matrix("H"), mpc.matrix("H")) and np.array_equal(mpc2.matrix("Fx"), mpc.matrix("Fx"))
    mpc.close(); mpc2.close()


@pytest.mark.parametrize("kernel", [4, 1])
def test_c3_quadrotor_batch_matches_oracle(quad, kernel):
    path, cfg, m = quad
    B = 48
    x0, xr = c3_batch(B, seed=0)
    mpc = sm.BatchedMimoMPC(path, batch=B, kernel=kernel, **EPS)
    # the oracle solves the QP the DEVICE assembled (assembly parity is the test above)
    H, A, ub = mpc.matrix("H"), mpc.matrix("A"), mpc.matrix("ub")
    q = oracle.mimo_batch_vectors(dict(Fx=mpc.matrix("Fx"), Fr=mpc.matrix("Fr")), x0, xr)
    ora = oracle.solve_batch(H, A, m["lb"], ub, q, np.tile(ub, (B, 1)), nthreads=os.cpu_count() or 1, **EPS)
    assert (ora["status"] == 1).all()
    mpc.set_state(x0=x0, xr=xr)
    assert mpc.controllerStep()
    x, y = mpc.solver.solution()
    info = mpc.solver.info()
    assert np.array_equal(info["status"], ora["status"])
    assert np.array_equal(info["iter"], ora["iter"])
    scale = np.abs(ora["x"]).max(axis=1)
    assert (np.abs(x - ora["x"]).max(axis=1) / scale).max() < REL
    u0 = mpc.control()
    assert (np.abs(u0 - ora["x"][:, :4]).max(axis=1) / np.abs(ora["x"][:, :4]).max(axis=1)).max() < REL
    # same active set: rows at their bound on one side are at their bound on the other
    act = lambda z: (np.tile(ub, (B, 1)) - z @ A.T) < 1e-6
    assert np.array_equal(act(x), act(ora["x"]))
    assert np.abs(y - ora["y"]).max() < 1e-6 * max(1.0, np.abs(ora["y"]).max())
    mpc.close()


def test_c3_kkt_properties_at_a_larger_batch(quad):
    """Size-independent properties (no oracle): stationarity, feasibility, complementarity of every instance,
    and shard equality -- a contiguous shard of the batch solved alone gives bit-identical results."""
    path, cfg, m = quad
    B = 2048
    x0, xr = c3_batch(B, seed=3)
    mpc = sm.BatchedMimoMPC(path, batch=B, **EPS)
    mpc.set_state(x0=x0, xr=xr)
    assert mpc.controllerStep()
    x, y = mpc.solver.solution()
    H, A, ub = mpc.matrix("H"), mpc.matrix("A"), mpc.matrix("ub")
    q = oracle.mimo_batch_vectors(dict(Fx=mpc.matrix("Fx"), Fr=mpc.matrix("Fr")), x0, xr)
    stat = np.abs(x @ H + q + y @ A).max(axis=1)
    assert (stat <= 1e-5 + 1e-5 * np.maximum(np.abs(x @ H).max(axis=1), np.abs(q).max(axis=1))).all()
    assert ((x @ A.T - ub).max(axis=1) <= 1e-5 + 1e-5 * np.abs(ub).max()).all()
    assert (y >= -1e-9).all()
    assert (np.abs(y * (ub - x @ A.T)).max(axis=1) <= 1e-4 * np.maximum(1.0, np.abs(y).max(axis=1))).all()
    lo, hi = 512, 1024
    shard = sm.BatchedMimoMPC(path, batch=hi - lo, **EPS)
    shard.set_state(x0=x0[lo:hi], xr=xr[lo:hi])
    assert shard.controllerStep()
    xs, _ = shard.solver.solution()
    assert np.array_equal(xs, x[lo:hi])
    mpc.close(); shard.close()


def test_c3_solution_is_optimal_for_the_rolled_out_plant(quad):
    """Independent of every assembled matrix (device or oracle): the inputs the GPU returns respect the box, and no feasible
    perturbation of them lowers the cost accumulated by rolling the plant out stage by stage."""
    path, cfg, _ = quad
    B, N, nu = 6, cfg["N"], cfg["Bd"].shape[1]
    x0, xr = c3_batch(B, seed=5)
    mpc = sm.BatchedMimoMPC(path, batch=B, **EPS)
    mpc.set_state(x0=x0, xr=xr)
    assert mpc.controllerStep()
    U, _ = mpc.solver.solution()
    mpc.close()
    lo, hi = np.tile(cfg["umin"], N), np.tile(cfg["umax"], N)

    def cost(b, z):
        x, J = x0[b].copy(), 0.0
        for k in range(N):
            u = z[k * nu:(k + 1) * nu]
            x = cfg["Ad"] @ x + cfg["Bd"] @ u
            J += ((x - xr[b]) ** 2 * cfg["Q"]).sum() + (u ** 2 * cfg["R"]).sum()
        return J
    rng = np.random.default_rng(1)
    for b in range(B):
        z = U[b]
        assert (z <= hi + 1e-4).all() and (z >= lo - 1e-4).all()
        zc = np.clip(z, lo, hi)
        J0 = cost(b, zc)
        for _ in range(12):
            zp = np.clip(zc + 10.0 ** rng.uniform(-4, -1) * rng.standard_normal(z.size), lo, hi)
            assert cost(b, zp) >= J0 - 1e-6 * max(1.0, abs(J0))
