"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs,
against the committed golden fixtures, and -- at BASELINE sizes -- through solver-independent KKT properties.

Tolerance (north star): u0 and the whole trajectory within 1e-4 relative (to ||x||_inf of the instance), same
status, same active set; additionally the iteration counts must agree because both sides run the same algorithm.
"""
import os

import numpy as np
import pytest

import oracle
import solvempc_b200 as sm
from problems import KNOWN_CASES, c2_batch, random_qp

pytestmark = pytest.mark.gpu

EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
REL = 1e-4          # north-star tolerance
TIGHT = 1e-7        # what two implementations of the same iteration actually achieve


def rel_err(x, xo):
    """max over instances of ||x - xo||_inf / ||xo||_inf; NaN solutions (infeasible instances) must coincide."""
    x, xo = np.asarray(x, dtype=float), np.asarray(xo, dtype=float)
    assert np.array_equal(np.isnan(x), np.isnan(xo)), "NaN pattern differs"
    x, xo = np.nan_to_num(x), np.nan_to_num(xo)
    scale = np.maximum(np.abs(xo).max(axis=-1), 1e-9)
    return (np.abs(x - xo).max(axis=-1) / scale).max()


def active_set(A, x, y, u, thr=1e-6):
    """a row is active if its dual is non-negligible or its slack is ~0."""
    Ax = x @ A.T
    return (y > thr * np.abs(y).max(axis=-1, keepdims=True) + 1e-9) | (u - Ax < thr * np.maximum(1.0, np.abs(u)))


def same_active_set(A, x, y, xo, yo, u):
    """Same active set, with hysteresis on the two thresholds so that rows sitting exactly at a threshold
    (duals ~1e-6 relative) cannot flip the comparison: clearly-active rows of one side must be active on the other."""
    lo_a, lo_b = active_set(A, x, y, u, 5e-7), active_set(A, xo, yo, u, 5e-7)
    hi_a, hi_b = active_set(A, x, y, u, 2e-6), active_set(A, xo, yo, u, 2e-6)
    return bool((hi_a <= lo_b).all() and (hi_b <= lo_a).all())


@pytest.fixture(scope="module")
def cfg_path(repo_root):
    return os.path.join(repo_root, "config", "MPC_API.json")


# (kernel 2 on a plan whose rows are [G; -G] pairs -- every MPC plan -- is the fused one-phase kernel when SMPC_SMALL_FUSED=1)
KERNEL_NAMES = {1: "admm_shared_generic_kernel",
                2: "admm_shared_small_fused_kernel" if os.environ.get("SMPC_SMALL_FUSED", "0") != "0" else "admm_shared_small_kernel", 4: "admm_shared_tile_kernel",
                5: "admm_shared_small_mma_kernel"}


@pytest.mark.parametrize("kernel", [1, 2, 4, 5])
def test_c1_reference_cases_through_mpc_api(cfg_path, golden, ref_mats, kernel):
    """Config 1: shipped plant + horizon, the four known-answer states, one controller each (B=1) and batched."""
    m, _ = ref_mats
    for batch in (1, 4):
        for k0 in range(0, 4, batch):
            cases = golden["cases"][k0:k0 + batch]
            mpc = sm.BatchedModelPredictiveControlAPI(cfg_path, batch=batch, kernel=kernel, **EPS)
            assert (mpc.mpcWindow, mpc.N_S, mpc.n_variables, mpc.n_constraints) == (15, 4, 15, 30)
            for name in ("H", "Gbar", "Fx", "Fu", "Fr", "Sbar", "Ku", "W0", "Sx", "Su", "CAB"):
                g = np.array(golden[name], dtype=float).reshape(mpc.matrix(name).shape)
                assert np.abs(mpc.matrix(name) - g).max() <= 1e-12 * max(np.abs(g).max(), 1e-300), name
            mpc.set_state(X=np.array([c["X"] for c in cases]), U=np.array([c["U"] for c in cases]),
                          ref=np.array([c["xref"] for c in cases]))
            assert mpc.controllerStep()
            f, ub = mpc.step_vectors()
            x, y = mpc.solver.solution()
            info = mpc.solver.info()
            _, U = mpc.state()
            for j, c in enumerate(cases):
                assert np.abs(f[j] - c["q"]).max() <= 1e-12 * np.abs(c["q"]).max()
                assert np.abs(ub[j] - c["u"]).max() <= 1e-12 * np.abs(c["u"]).max()
                assert info["status"][j] == c["status"] == 1 and info["iter"][j] == c["iter"]
                assert rel_err(x[j], np.array(c["x"])) < TIGHT
                assert abs(U[j] - c["U_after"]) < 1e-9
            mpc.close()
    # and against the exact KKT answers (solver independent)
    mpc = sm.BatchedModelPredictiveControlAPI(cfg_path, batch=4, kernel=kernel, **EPS)
    mpc.set_state(X=np.array([c["X"] for c in KNOWN_CASES]), U=np.array([c["U"] for c in KNOWN_CASES]),
                  ref=np.array([c["ref"] for c in KNOWN_CASES]))
    assert mpc.controllerStep()
    x, y = mpc.solver.solution()
    f, ub = mpc.step_vectors()
    for j, c in enumerate(KNOWN_CASES):
        xs, ys = oracle.exact_qp_active_set(m["H"], f[j], m["Gbar"], ub[j], c["active"])
        assert np.abs(x[j] - xs).max() < 5e-6 and abs(x[j][0] - c["du0"]) < 1e-6
        assert set(np.nonzero(active_set(m["Gbar"], x[j], y[j], ub[j]))[0]) == set(c["active"])


@pytest.mark.parametrize("kernel", [1, 2, 4, 5])
def test_c2_batch_4096_matches_oracle(ref_mats, kernel):
    """Config 2: 4096 random x0 / references sharing P and A, cold solves."""
    m, _ = ref_mats
    B = 4096
    X, U, ref = c2_batch(B, seed=0)
    f, ub = oracle.mpc_batch_vectors(m, X, U, ref)
    ora = oracle.solve_batch(m["H"], m["Gbar"], m["lb"], m["W0"], f, ub, nthreads=os.cpu_count() or 1, **EPS)
    s = sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=B, kernel=kernel, **EPS)
    assert s.kernel_name == KERNEL_NAMES[kernel]
    s.update_gradient(f)
    s.update_upper_bound(ub)
    s.solve()
    x, y = s.solution()
    info = s.info()
    assert np.array_equal(info["status"], ora["status"]) and (info["status"] == 1).all()
    assert np.array_equal(info["iter"], ora["iter"])
    assert rel_err(x, ora["x"]) < TIGHT < REL
    assert rel_err(x[:, :1], ora["x"][:, :1]) < REL        # u0 increment
    assert np.abs(y - ora["y"]).max() < 1e-7 * max(1.0, np.abs(ora["y"]).max())
    assert same_active_set(m["Gbar"], x, y, ora["x"], ora["y"], ub)
    assert s.count_solved() == B
    # second solve on the same handle is warm started and rho persists (cpp:52): oracle does the same per instance
    s.solve()
    info2 = s.info()
    assert (info2["iter"] == 25).all() and (info2["status"] == 1).all()
    # cold-solve mode reproduces the first solve bit for bit
    s.set_cold_solves(True)
    s.solve()
    x3, _ = s.solution()
    assert np.array_equal(x3, x) and np.array_equal(s.info()["iter"], info["iter"])
    s.close()


def test_generic_and_small_kernels_agree_bitwise_on_status(ref_mats):
    m, _ = ref_mats
    X, U, ref = c2_batch(512, seed=5)
    f, ub = oracle.mpc_batch_vectors(m, X, U, ref)
    out = []
    for kernel in (1, 2, 4):
        s = sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=512, kernel=kernel, **EPS)
        s.update_gradient(f); s.update_upper_bound(ub); s.solve()
        out.append((s.solution()[0], s.info()))
        s.close()
    assert np.array_equal(out[0][1]["iter"], out[1][1]["iter"]) and np.array_equal(out[0][1]["iter"], out[2][1]["iter"])
    assert rel_err(out[0][0], out[1][0]) < 1e-9 and rel_err(out[0][0], out[2][0]) < 1e-9
    # the longest-expected-first schedule of the small kernel changes the order of work only: bitwise same answers
    s = sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=512, kernel=2, **EPS)
    s.set_scheduling(False)
    s.update_gradient(f); s.update_upper_bound(ub); s.solve()
    x_off, y_off = s.solution()
    assert np.array_equal(x_off, out[1][0]) and np.array_equal(s.info()["iter"], out[1][1]["iter"])
    s.close()


@pytest.mark.parametrize("kernel", [2, 5])
@pytest.mark.parametrize("n,m,B", [(3, 0, 8), (12, 20, 64), (16, 32, 64), (9, 14, 19)])
def test_random_qps_small_kernels(n, m, B, kernel):
    """The two small-QP kernels (one warp per QP; DMMA tile of 8 QPs) on every shape class they accept."""
    _random_qps(n, m, B, kernel)


@pytest.mark.parametrize("kernel", [1, 4])
@pytest.mark.parametrize("n,m,B", [(3, 0, 8), (12, 20, 64), (16, 32, 64), (17, 33, 32), (40, 70, 37), (100, 200, 8)])
def test_random_qps_generic_shapes(n, m, B, kernel):
    _random_qps(n, m, B, kernel)


def _random_qps(n, m, B, kernel):
    """Two-sided bounds, equality rows and a free row; every instance gets its own q, l, u.  The tile kernel (4)
    sums its dot products in DMMA order, so an instance sitting on a termination threshold may stop one check
    (25 iterations) away from the oracle: iteration counts must agree on >= 90 % of the instances and the
    solutions within the north-star tolerance everywhere."""
    P, q0, A, l0, u0 = random_qp(n, max(m, 1), seed=n * 7 + m)
    if m == 0:
        A, l0, u0 = np.zeros((0, n)), np.zeros(0), np.zeros(0)
    else:
        l0[: m // 5] = u0[: m // 5]
        l0[m // 5], u0[m // 5] = -np.inf, np.inf
    rng = np.random.default_rng(n + m)
    q = q0[None, :] + 0.3 * rng.standard_normal((B, n))
    sh = 0.2 * rng.standard_normal((B, m)) if m else np.zeros((B, 0))
    l, u = l0[None, :] + sh, u0[None, :] + sh
    xs, ys, st, it = [], [], [], []
    for b in range(B):
        so = oracle.Solver(P, np.zeros(n), A, l0, u0, **EPS)
        so.update_lin_cost(q[b])
        if m:
            so.update_bounds(l[b], u[b])
        r = so.solve()
        xs.append(r["x"]); ys.append(r["y"]); st.append(r["status"]); it.append(r["iter"])
    s = sm.BatchedSolver(P, A, l0, u0, batch=B, kernel=kernel, **EPS)
    assert s.kernel_name == ("admm_shared_small_kernel" if kernel == 2 else KERNEL_NAMES[kernel])   # rows are not pairs here
    s.update_gradient(q)
    if m:
        s.update_bounds(l, u)
    s.solve()
    x, y = s.solution()
    info = s.info()
    assert np.array_equal(info["status"], np.array(st))
    same = info["iter"] == np.array(it)
    if kernel == 1:
        assert same.all()
    assert same.mean() >= 0.9 and rel_err(x, np.array(xs)) < REL
    assert rel_err(x[same], np.array(xs)[same]) < TIGHT
    if m:
        assert rel_err(y[same], np.array(ys)[same]) < 1e-5
    assert (np.array(st) == 1).sum() >= 1          # (the shifted bounds make some instances infeasible: NaN parity above)
    s.close()


@pytest.mark.parametrize("n,mp,B", [(3, 2, 9), (10, 10, 64), (16, 16, 64), (15, 15, 48), (12, 7, 33)])
def test_small_kernel_row_pairs(n, mp, B):
    """A = [G; -G] (the reference's two-sided limit, cpp:335): the one-warp kernel multiplies the top half of W only.
    Pairs below, at and above the lane-mapping edges (mp < 15, mp = 16), per-instance bounds that make some pairs
    inconsistent (u_top + u_bot < 0: primal infeasible) and +inf upper bounds on a few rows."""
    P, q0, G, _, _ = random_qp(n, mp, seed=31 * n + mp)
    A = np.vstack([G, -G])
    m = 2 * mp
    rng = np.random.default_rng(n * mp)
    u0 = np.concatenate([1.0 + rng.random(mp), 1.0 + rng.random(mp)])
    l0 = np.full(m, -np.inf)
    q = q0[None, :] + 0.5 * rng.standard_normal((B, n))
    u = u0[None, :] + 0.4 * rng.standard_normal((B, m))
    u[1::7, 0] = -2.0; u[1::7, mp] = -2.0          # G x <= -2 and -G x <= -2: infeasible pair
    s = sm.BatchedSolver(P, A, l0, u0, batch=B, kernel=2, **EPS)
    assert s.kernel_name == KERNEL_NAMES[2] and s.row_pairs == mp
    s.update_gradient(q); s.update_upper_bound(u); s.solve()
    x, y = s.solution(); info = s.info()
    ora = oracle.solve_batch(P, A, l0, u0, q, u, nthreads=os.cpu_count() or 1, **EPS)
    assert np.array_equal(info["status"], ora["status"]) and np.array_equal(info["iter"], ora["iter"])
    ok = ora["status"] == 1
    assert ok.sum() >= B // 2 and (ora["status"][1::7] == sm.PRIMAL_INFEASIBLE).all()
    assert rel_err(x[ok], ora["x"][ok]) < TIGHT and rel_err(y[ok], ora["y"][ok]) < 1e-5
    assert np.isnan(x[~ok]).all()
    # generic kernel (no pair exploitation) on the same data: same statuses and iteration counts
    g = sm.BatchedSolver(P, A, l0, u0, batch=B, kernel=1, **EPS)
    g.update_gradient(q); g.update_upper_bound(u); g.solve()
    gi = g.info()
    assert g.row_pairs == 0 and np.array_equal(gi["status"], info["status"]) and np.array_equal(gi["iter"], info["iter"])
    s.close(); g.close()


@pytest.mark.parametrize("kernel", [1, 2, 4, 5])
@pytest.mark.parametrize("opts", [dict(adaptive_rho=0), dict(scaling=0), dict(scaled_termination=1),
                                  dict(adaptive_rho_interval=50), dict(check_termination=10, adaptive_rho_interval=30),
                                  dict(max_iter=30), dict(alpha=1.0, rho=1.0, eps_abs=1e-3, eps_rel=1e-3)])
def test_settings_follow_the_oracle(ref_mats, kernel, opts):
    m, _ = ref_mats
    kw = {**EPS, **opts}
    X, U, ref = c2_batch(96, seed=11)
    f, ub = oracle.mpc_batch_vectors(m, X, U, ref)
    ora = oracle.solve_batch(m["H"], m["Gbar"], m["lb"], m["W0"], f, ub, nthreads=4, **kw)
    s = sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=96, kernel=kernel, **kw)
    s.update_gradient(f); s.update_upper_bound(ub); s.solve()
    x, y = s.solution()
    info = s.info()
    assert np.array_equal(info["status"], ora["status"]) and np.array_equal(info["iter"], ora["iter"])
    assert rel_err(x, ora["x"]) < 1e-6
    s.close()


@pytest.mark.parametrize("kernel", [1, 2, 4, 5])
def test_infeasible_unbounded_and_bad_bounds(kernel):
    P = np.eye(2); A = np.array([[1.0, 0.0], [1.0, 0.0]])
    l = np.array([[1.0, -np.inf], [-2.0, -np.inf], [0.5, -np.inf]]); u = np.array([[np.inf, -1.0], [np.inf, 3.0], [np.inf, 0.25]])
    s = sm.BatchedSolver(P, A, np.array([-1.0, -np.inf]), np.array([np.inf, 1.0]), batch=3, kernel=kernel)
    s.update_gradient(np.zeros((3, 2))); s.update_bounds(l, u); s.solve()
    x, _ = s.solution(); info = s.info()
    exp = []
    for b in range(3):
        so = oracle.Solver(P, np.zeros(2), A, np.array([-1.0, -np.inf]), np.array([np.inf, 1.0]))
        so.update_bounds(l[b], u[b]); exp.append(so.solve())
    assert [e["status"] for e in exp] == [sm.PRIMAL_INFEASIBLE, sm.SOLVED, sm.PRIMAL_INFEASIBLE]
    assert list(info["status"]) == [e["status"] for e in exp] and list(info["iter"]) == [e["iter"] for e in exp]
    assert np.isnan(x[0]).all() and np.isnan(x[2]).all() and np.abs(x[1] - exp[1]["x"]).max() < 1e-9
    s.close()
    P = np.diag([1.0, 0.0]); A = np.array([[1.0, 0.0]])
    s = sm.BatchedSolver(P, A, np.array([-1.0]), np.array([1.0]), batch=2, kernel=kernel)
    s.update_gradient(np.array([[0.0, -1.0], [0.3, 0.0]])); s.solve()
    info = s.info(); x, _ = s.solution()
    so = oracle.Solver(P, np.zeros(2), A, np.array([-1.0]), np.array([1.0])); so.update_lin_cost(np.array([0.0, -1.0]))
    r0 = so.solve()
    assert r0["status"] == sm.DUAL_INFEASIBLE and info["status"][0] == sm.DUAL_INFEASIBLE and info["iter"][0] == r0["iter"]
    assert np.isnan(x[0]).all() and info["status"][1] == sm.SOLVED
    s.close()
    # l > u for one instance: that instance stays UNSOLVED, the others are solved
    P, q, A, l0, u0 = random_qp(6, 8, 2)
    s = sm.BatchedSolver(P, A, l0, u0, batch=2, kernel=kernel)
    lb = np.stack([l0, l0]); ubd = np.stack([u0, u0]); lb[1, 3] = ubd[1, 3] + 1.0
    s.update_gradient(np.stack([q, q])); s.update_bounds(lb, ubd); s.solve()
    info = s.info()
    assert info["status"][0] == sm.SOLVED and info["status"][1] == sm.UNSOLVED and info["iter"][1] == 0
    s.close()


@pytest.mark.parametrize("kernel", [1, 2, 4, 5])
def test_closed_loop_warm_start_matches_reference_run(cfg_path, golden, ref_mats, kernel):
    """40 warm-started controllerStep + plant steps: the golden trajectory came from the reference's own class."""
    m, cfg = ref_mats
    cl = golden["closed_loop"]
    B = 5
    rng = np.random.default_rng(4)
    X0 = np.vstack([cl["X0"], rng.standard_normal((B - 1, 4)) * np.array([0.05, 0.2, 0.05, 0.3])])
    U0 = np.r_[cl["U0"], rng.standard_normal(B - 1)]
    ref = np.r_[cl["xref"], 0.2 * rng.standard_normal(B - 1)]
    mpc = sm.BatchedModelPredictiveControlAPI(cfg_path, batch=B, kernel=kernel, **EPS)
    mpc.set_state(X=X0, U=U0, ref=ref)
    solvers = [oracle.Solver(m["H"], np.zeros(15), m["Gbar"], m["lb"], m["W0"], **EPS) for _ in range(B)]
    Xo, Uo = X0.copy(), U0.copy()
    for k in range(cl["steps"]):
        assert mpc.controllerStep()
        mpc.plant_step()
        X, U = mpc.state()
        info = mpc.solver.info()
        for b in range(B):
            f, ub = oracle.mpc_step_vectors(m, Xo[b], Uo[b], ref[b])
            solvers[b].update_lin_cost(f); solvers[b].update_upper_bound(ub)
            r = solvers[b].solve()
            Uo[b] += r["x"][0]
            Xo[b] = cfg["Ad"] @ Xo[b] + cfg["Bd"] * Uo[b]
            assert info["iter"][b] == r["iter"] and info["status"][b] == 1
        assert info["iter"][0] == cl["iters"][k]
        assert abs(U[0] - cl["U"][k]) < 1e-8 and np.abs(X[0] - np.array(cl["X"][k])).max() < 1e-8
        assert np.abs(U - Uo).max() < 1e-7 * max(1.0, np.abs(Uo).max())
        assert np.abs(X - Xo).max() < 1e-7 * max(1.0, np.abs(Xo).max())
    mpc.close()


def test_user_warm_start_and_device_buffers(ref_mats):
    import torch
    m, _ = ref_mats
    B = 64
    X, U, ref = c2_batch(B, seed=21)
    f, ub = oracle.mpc_batch_vectors(m, X, U, ref)
    s = sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=B, **EPS)
    fd, ud = torch.from_numpy(f).cuda(), torch.from_numpy(ub).cuda()
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    s.update_gradient(fd); s.update_upper_bound(ud); s.solve()
    xd = torch.empty(B, 15, dtype=torch.float64, device="cuda"); yd = torch.empty(B, 30, dtype=torch.float64, device="cuda")
    s.solution_into(xd, yd)
    torch.cuda.synchronize()
    x, y = s.solution()
    assert np.array_equal(xd.cpu().numpy(), x)
    it0 = s.info()["iter"].copy()
    # osqp_warm_start with the solution on a fresh handle: same path as the oracle's warm start
    s2 = sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=B, **EPS)
    s2.update_gradient(f); s2.update_upper_bound(ub); s2.warm_start(x, y); s2.solve()
    it2 = s2.info()["iter"]
    x2, _ = s2.solution()
    for b in range(0, B, 7):
        so = oracle.Solver(m["H"], np.zeros(15), m["Gbar"], m["lb"], m["W0"], **EPS)
        so.update_lin_cost(f[b]); so.update_upper_bound(ub[b]); so.warm_start(x[b], y[b])
        r = so.solve()
        assert r["iter"] == it2[b] and rel_err(x2[b], r["x"]) < 1e-6
    assert (it2 <= it0).all()
    s.close(); s2.close()


def test_full_size_properties_and_sharding(ref_mats):
    """BASELINE-size batch (65 536 instances) checked through size-independent properties: every instance
    satisfies OSQP's stopping rule recomputed independently in unscaled space, iterations are multiples of 25,
    and a shard of the batch solved on its own handle gives bitwise the same answers."""
    m, _ = ref_mats
    B = 65536
    X, U, ref = c2_batch(B, seed=123)
    f, ub = oracle.mpc_batch_vectors(m, X, U, ref)
    s = sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=B, **EPS)
    s.update_gradient(f); s.update_upper_bound(ub); s.solve()
    x, y = s.solution(); info = s.info()
    assert (info["status"] == 1).all() and (info["iter"] % 25 == 0).all() and info["iter"].max() <= 400
    H, G = m["H"], m["Gbar"]
    Ax, Px, Aty = x @ G.T, x @ H.T, y @ G
    rd = np.abs(Px + f + Aty).max(axis=1)
    tol_d = 1e-5 + 1e-5 * np.maximum(np.maximum(np.abs(Px).max(1), np.abs(Aty).max(1)), np.abs(f).max(1))
    assert (rd < tol_d * 1.0000001).all()
    assert ((Ax - ub).max(axis=1) < 1e-5 + 1e-5 * np.abs(Ax).max(axis=1)).all()
    assert (y > -1e-9).all()                                   # one-sided rows: duals >= 0
    tol_p = (1e-5 + 1e-5 * np.abs(Ax).max(axis=1))[:, None]
    slack = ub - Ax
    assert (slack[y > 1e-12] <= (3 * tol_p * np.ones_like(y))[y > 1e-12]).all()      # y_i > 0  =>  row at its bound
    assert (y[slack > 3 * tol_p * np.ones_like(y)] <= 1e-12).all()                   # slack   =>  y_i = 0
    assert 0.3 < (np.abs(y).max(axis=1) > 1e-7).mean() < 0.6   # ~43 % of instances saturate
    lo, hi = 3 * B // 8, 5 * B // 8
    s2 = sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=hi - lo, **EPS)
    s2.update_gradient(f[lo:hi]); s2.update_upper_bound(ub[lo:hi]); s2.solve()
    x2, y2 = s2.solution()
    assert np.array_equal(x2, x[lo:hi]) and np.array_equal(y2, y[lo:hi])
    # a 2 048-instance sample against the oracle
    idx = np.arange(0, B, 32)
    ora = oracle.solve_batch(m["H"], m["Gbar"], m["lb"], m["W0"], f[idx], ub[idx], nthreads=os.cpu_count() or 1, **EPS)
    assert np.array_equal(info["iter"][idx], ora["iter"]) and rel_err(x[idx], ora["x"]) < TIGHT
    s.close(); s2.close()


@pytest.mark.parametrize("kernel,B", [(1, 16), (4, 16), (4, 300)])
def test_horizon_100_config5_shape(ref_mats, repo_root, kernel, B):
    """Config 5 shape (N = 100 -> n = 100, m = 200) through the MPC API on the generic and the DMMA tile kernel
    (B = 300 > one wave of 16-slot tiles per SM exercises the slot recycling)."""
    _, cfg = ref_mats
    N = 100
    mats = oracle.mpc_build(**{**cfg, "N": N})
    conf = dict(Ad=cfg["Ad"], Bd=cfg["Bd"], Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=N)
    mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=B, kernel=kernel, **EPS)
    assert mpc.solver.kernel_name == KERNEL_NAMES[kernel]
    for name in ("H", "Gbar", "Fx", "Fu", "Fr", "Sbar"):
        assert np.abs(mpc.matrix(name) - mats[name]).max() <= 1e-11 * np.abs(mats[name]).max(), name
    X, U, ref = c2_batch(B, seed=9)
    mpc.set_state(X=X, U=U, ref=ref)
    ok = mpc.controllerStep()
    f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
    fd, ubd = mpc.step_vectors()
    assert np.abs(fd - f).max() < 1e-11 * np.abs(f).max() and np.abs(ubd - ub).max() < 1e-11 * np.abs(ub).max()
    ora = oracle.solve_batch(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], f, ub, nthreads=os.cpu_count() or 1, **EPS)
    x, _ = mpc.solver.solution(); info = mpc.solver.info()
    assert ok == bool((ora["status"] == 1).all())
    assert np.array_equal(info["status"], ora["status"]) and np.array_equal(info["iter"], ora["iter"])
    assert rel_err(x, ora["x"]) < 1e-6
    mpc.close()


def test_control_and_status_in_one_transfer(cfg_path):
    """smpc_mpc_get_control_status = what the reference's main loop reads after controllerStep (solver.cpp:55-60)."""
    B = 33
    X, U, ref = c2_batch(B, seed=9)
    mpc = sm.BatchedModelPredictiveControlAPI(cfg_path, batch=B, **EPS)
    mpc.set_state(X=X, U=U, ref=ref)
    mpc.controller_step_async()
    Uo, st = np.empty(B), np.empty(B, np.int32)
    mpc.results_into(Uo, st)
    _, U2 = mpc.state()
    assert np.array_equal(Uo, U2) and np.array_equal(st, mpc.solver.info()["status"]) and (st == 1).all()
    x, _ = mpc.solver.solution()
    assert np.array_equal(Uo, U + x[:, 0])
    mpc.close()


@pytest.mark.parametrize("kernel", [2, 4])
def test_controller_step_from_equals_set_state_plus_step(cfg_path, kernel):
    """smpc_mpc_controller_step_from = set_state + controllerStep (src/solver.cpp:45-55) in one call: same bits from device
    tensors, pinned host tensors (both read by the step's first kernel on the small-QP path) and pageable host arrays
    (two-call path), and the controller's own X / U follow."""
    import torch
    B = 300
    X, U, ref = c2_batch(B, seed=33)
    base = sm.BatchedModelPredictiveControlAPI(cfg_path, batch=B, kernel=kernel, **EPS)
    base.set_state(X=X, U=U, ref=ref)
    base.controller_step_async()
    x0, _ = base.solver.solution(); X0, U0 = base.state(); st0 = base.solver.info()["status"]
    assert (st0 == 1).all() and np.array_equal(U0, U + x0[:, 0])
    for how in ("device", "pinned", "pageable"):
        mpc = sm.BatchedModelPredictiveControlAPI(cfg_path, batch=B, kernel=kernel, **EPS)
        if how == "device":
            args = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (X, U, ref)]
        elif how == "pinned":
            args = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (X, U, ref)]
        else:
            args = [np.ascontiguousarray(a) for a in (X, U, ref)]
        l0 = mpc.launches
        mpc.controller_step_from(*args)
        if kernel == 2 and how != "pageable":
            assert mpc.launches == l0 + 2               # step vectors + pre-pass, ADMM: no gather kernel in front
        x1, _ = mpc.solver.solution(); X1, U1 = mpc.state()
        assert np.array_equal(x1, x0) and np.array_equal(U1, U0) and np.array_equal(X1, X), how
        assert np.array_equal(mpc.solver.info()["iter"], base.solver.info()["iter"])
        # a second step from the controller's own state (the copies written by the fused kernel)
        mpc.controller_step_async(); base_again = None
        mpc.close()
    with pytest.raises(Exception):
        base.controller_step_from(X, None, ref)
    base.close()


@pytest.mark.parametrize("kernel", [2, 4])
def test_bound_results_follow_every_step(cfg_path, kernel):
    """smpc_mpc_bind_results: U and the statuses land in the caller's pinned (or device) buffers with the step itself -- stored by
    the one-warp kernel as each instance ends, by one export launch behind the other kernels -- and equal what
    get_control_status returns; pageable host memory is refused."""
    import torch
    B = 513
    X, U, ref = c2_batch(B, seed=5)
    mpc = sm.BatchedModelPredictiveControlAPI(cfg_path, batch=B, kernel=kernel, **EPS)
    Uo, st = torch.zeros(B, dtype=torch.float64).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()
    mpc.bind_results(Uo, st)
    mpc.controller_step_from(X, U, ref)
    mpc.sync()
    U1, s1 = np.empty(B), np.empty(B, np.int32)
    mpc.results_into(U1, s1)
    x, _ = mpc.solver.solution()
    assert np.array_equal(Uo.numpy(), U1) and np.array_equal(st.numpy(), s1) and (s1 == 1).all()
    assert np.array_equal(U1, U + x[:, 0])
    # a state far outside the operating range: whatever the status, the buffers still equal the controller's own U / status
    Xb = X.copy(); Xb[7] = [50.0, 0.0, 50.0, 0.0]
    mpc.controller_step_from(Xb, U, ref)
    mpc.sync()
    mpc.results_into(U1, s1)
    assert np.array_equal(Uo.numpy(), U1) and np.array_equal(st.numpy(), s1)
    if s1[7] != 1:
        assert U1[7] == U[7]
    dU, ds = torch.zeros(B, dtype=torch.float64, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda")
    mpc.bind_results(dU, ds)
    mpc.controller_step_from(X, U, ref); mpc.sync()
    mpc.results_into(U1, s1)
    assert np.array_equal(dU.cpu().numpy(), U1) and np.array_equal(ds.cpu().numpy(), s1)
    with pytest.raises(Exception):
        mpc.bind_results(np.zeros(B), np.zeros(B, np.int32))          # pageable host memory
    mpc.bind_results(None, None)
    Uo.zero_()
    mpc.controller_step_from(X, U, ref); mpc.sync()
    assert (Uo.numpy() == 0).all()                                     # unbound
    mpc.close()


def test_pinned_host_buffers_take_the_zero_copy_path(cfg_path):
    """Pinned host buffers are gathered / exported by a kernel over PCIe, pageable ones by cudaMemcpyAsync: same results."""
    import torch
    B = 257
    X, U, ref = c2_batch(B, seed=21)
    out = []
    for pinned in (False, True):
        mpc = sm.BatchedModelPredictiveControlAPI(cfg_path, batch=B, **EPS)
        if pinned:
            hx, hu, hr = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (X, U, ref)]
            Uo, st = torch.empty(B, dtype=torch.float64).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory()
            l0 = mpc.launches
            mpc.set_state(X=hx, U=hu, ref=hr)
            assert mpc.launches == l0 + 1                      # one gather kernel, no copy nodes
            mpc.controller_step_async()
            mpc.results_into(Uo, st)
            out.append((Uo.numpy().copy(), st.numpy().copy()))
            mpc.set_state(U=hu)                                # partial updates work too
            mpc.results_into(Uo, st)
            assert np.array_equal(Uo.numpy(), U)
        else:
            mpc.set_state(X=X, U=U, ref=ref)
            mpc.controller_step_async()
            Uo, st = np.empty(B), np.empty(B, np.int32)
            mpc.results_into(Uo, st)
            out.append((Uo, st))
        mpc.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and (out[0][1] == 1).all()


def test_fused_small_kernel_opt_in_matches_oracle(repo_root):
    """The opt-in one-phase small-QP kernel (SMPC_SMALL_FUSED=1, read once per process: hence the subprocess): statuses and
    iteration counts of the oracle on config-2 instances, cold and warm, and on a plan with infeasible pairs."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import os, sys
        import numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
        import oracle, solvempc_b200 as sm
        from problems import c2_batch, random_qp
        EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
        cfg = oracle.load_config(os.path.join(%r, "config", "MPC_API.json"))
        m = oracle.mpc_build(**cfg)
        B = 777
        X, U, ref = c2_batch(B, seed=11)
        f, ub = oracle.mpc_batch_vectors(m, X, U, ref)
        ora = oracle.solve_batch(m["H"], m["Gbar"], m["lb"], m["W0"], f, ub, nthreads=os.cpu_count() or 1, **EPS)
        s = sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=B, kernel=2, **EPS)
        assert s.kernel_name == "admm_shared_small_fused_kernel", s.kernel_name
        s.update_gradient(f); s.update_upper_bound(ub); s.solve()
        x, y = s.solution(); info = s.info()
        assert np.array_equal(info["status"], ora["status"]) and np.array_equal(info["iter"], ora["iter"])
        sc = np.maximum(np.abs(ora["x"]).max(axis=1), 1e-9)
        assert (np.abs(x - ora["x"]).max(axis=1) / sc).max() < 1e-9
        s.solve()                                     # warm: converged iterates, first check ends every solve
        assert (s.info()["iter"] == 25).all() and (s.info()["status"] == 1).all()
        s.close()
        # random pairs [G; -G] with some infeasible instances
        n, mp = 9, 13
        P, q0, G, _, _ = random_qp(n, mp, seed=5)
        A = np.vstack([G, -G]); rng = np.random.default_rng(3)
        u0 = np.concatenate([1.0 + rng.random(mp), 1.0 + rng.random(mp)]); l0 = np.full(2 * mp, -np.inf)
        q = q0[None, :] + 0.5 * rng.standard_normal((64, n)); u = u0[None, :] + 0.4 * rng.standard_normal((64, 2 * mp))
        u[1::7, 0] = -2.0; u[1::7, mp] = -2.0
        s = sm.BatchedSolver(P, A, l0, u0, batch=64, kernel=2, **EPS)
        s.update_gradient(q); s.update_upper_bound(u); s.solve()
        info = s.info()
        ora = oracle.solve_batch(P, A, l0, u0, q, u, nthreads=4, **EPS)
        assert np.array_equal(info["status"], ora["status"]) and np.array_equal(info["iter"], ora["iter"])
        assert (info["status"] == -3).sum() == len(range(1, 64, 7))
        s.close()
        print("fused ok")
    """) % (repo_root, repo_root, repo_root)
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SMPC_SMALL_FUSED="1"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "fused ok" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


@pytest.mark.parametrize("kernel", [1, 2, 4, 5])
def test_row_class_change_is_solved_with_the_plans_rho_vec(kernel):
    """Shared-factor regime: instances whose own bounds turn an inequality row into an equality or a free row (OSQP would
    re-classify and refactor for them) are solved with the setup's rho_vec entries: same status and solution as the oracle within
    the north star's tolerance; the iteration count may differ for those instances, and must not for the untouched ones."""
    n, m, B = 8, 12, 6
    P, q, A, l0, u0 = random_qp(n, m, 21)
    s = sm.BatchedSolver(P, A, l0, u0, batch=B, kernel=kernel, **EPS)
    l, u = np.tile(l0, (B, 1)), np.tile(u0, (B, 1))
    mid = 0.5 * (l0 + u0)
    l[1, 2] = u[1, 2] = mid[2]                       # equality row
    l[2, 5], u[2, 5] = -np.inf, np.inf               # free row
    l[3, 0] = u[3, 0] = mid[0]; l[3, 7], u[3, 7] = -np.inf, np.inf
    qq = np.tile(q, (B, 1)) + 0.1 * np.arange(B)[:, None]
    s.update_gradient(qq); s.update_bounds(l, u); s.solve()
    x, _ = s.solution(); info = s.info()
    s.close()
    for b in range(B):
        so = oracle.Solver(P, np.zeros(n), A, l0, u0, **EPS)
        so.update_lin_cost(qq[b]); so.update_bounds(l[b], u[b])
        r = so.solve()
        assert info["status"][b] == r["status"] == sm.SOLVED
        assert np.abs(x[b] - r["x"]).max() <= 1e-4 * max(1.0, np.abs(r["x"]).max())
        if b in (0, 4, 5):
            assert info["iter"][b] == r["iter"] and np.abs(x[b] - r["x"]).max() <= 1e-9 * max(1.0, np.abs(r["x"]).max())
        if b in (1, 3):
            assert abs((A @ x[b])[2 if b == 1 else 0] - (mid[2] if b == 1 else mid[0])) < 1e-3


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_c2_more_seeds_match_the_oracle(ref_mats, seed):
    """Config 2 on other seeded batches (the default small-QP kernel): statuses and iteration counts of the oracle, instance by
    instance -- a soak for the round-off-level liberties of the kernel's check (bit-pattern norms, reciprocal-based rho estimate)."""
    m, _ = ref_mats
    B = 4096
    X, U, ref = c2_batch(B, seed=seed)
    f, ub = oracle.mpc_batch_vectors(m, X, U, ref)
    ora = oracle.solve_batch(m["H"], m["Gbar"], m["lb"], m["W0"], f, ub, nthreads=os.cpu_count() or 1, **EPS)
    s = sm.BatchedSolver(m["H"], m["Gbar"], m["lb"], m["W0"], batch=B, kernel=2, **EPS)
    s.update_gradient(f); s.update_upper_bound(ub); s.solve()
    x, _ = s.solution(); info = s.info()
    s.close()
    assert np.array_equal(info["status"], ora["status"]) and np.array_equal(info["iter"], ora["iter"])
    sc = np.maximum(np.abs(ora["x"]).max(axis=1), 1e-12)
    assert (np.abs(x - ora["x"]).max(axis=1) / sc).max() < 1e-9


@pytest.mark.parametrize("N", [50, 64])
def test_tile_kernel_two_tiles_per_sm_matches_the_oracle(ref_mats, N):
    """Horizons whose tiles share an SM pairwise (admm_shared_tile_kernel, CTAS = 2 variant): 2 048 cold MPC QPs, every status and
    iteration count as the oracle's."""
    _, cfg = ref_mats
    B = 2048
    mats = oracle.mpc_build(**{**cfg, "N": N})
    X, U, ref = c2_batch(B, seed=40 + N)
    f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
    s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=B, kernel=4, **EPS)
    s.update_gradient(f); s.update_upper_bound(ub); s.solve()
    x, _ = s.solution(); info = s.info()
    s.close()
    ora = oracle.solve_batch(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], f, ub, nthreads=os.cpu_count() or 1, **EPS)
    assert np.array_equal(info["status"], ora["status"]) and np.array_equal(info["iter"], ora["iter"])
    assert (np.abs(x - ora["x"]).max(axis=1) / np.maximum(np.abs(ora["x"]).max(axis=1), 1e-12)).max() < 1e-8
