"""Development driver (GPU): tile kernel (kernel=4) against the generic kernel (kernel=1) and timing of both."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import solvempc_b200 as sm
from problems import c2_batch, random_qp

EPS = dict(eps_abs=1e-5, eps_rel=1e-5)

def run(P, A, l0, u0, q, l, u, kernel, reps=1, **kw):
    B = q.shape[0]
    s = sm.BatchedSolver(P, A, l0, u0, batch=B, kernel=kernel, **{**EPS, **kw})
    s.update_gradient(q)
    if l is not None: s.update_bounds(l, u)
    s.set_cold_solves(True)
    s.solve(); s.sync()
    s.enable_timing(True); s.kernel_ms(reset=True)
    for _ in range(reps): s.solve()
    s.sync()
    ms, cnt = s.kernel_ms()
    x, y = s.solution(); info = s.info()
    name = s.kernel_name
    s.close()
    return x, y, info, ms / max(cnt, 1), name

def compare(tag, a, b):
    xa, ya, ia, ta, na = a; xb, yb, ib, tb, nb = b
    same_st = (ia["status"] == ib["status"]).mean()
    same_it = (ia["iter"] == ib["iter"]).mean()
    nan_eq = np.array_equal(np.isnan(xa), np.isnan(xb))
    xa, xb = np.nan_to_num(xa), np.nan_to_num(xb)
    sc = np.maximum(np.abs(xb).max(axis=1), 1e-9)
    err = np.abs(xa - xb).max(axis=1) / sc
    eq = ia["iter"] == ib["iter"]
    print(f"{tag}: status eq {same_st:.4f} iter eq {same_it:.4f} nan_eq {nan_eq} relerr(all) {err.max():.2e} relerr(iter-eq) {err[eq].max() if eq.any() else -1:.2e} "
          f"iters mean {ib['iter'].mean():.1f} max {ib['iter'].max()} | {na} {ta:.3f} ms vs {nb} {tb:.3f} ms", flush=True)

def shape_case(n, m, B, seed, reps=1):
    P, q0, A, l0, u0 = random_qp(n, m, seed=seed)
    l0[: m // 5] = u0[: m // 5]
    l0[m // 5], u0[m // 5] = -np.inf, np.inf
    rng = np.random.default_rng(n + m)
    q = q0[None, :] + 0.3 * rng.standard_normal((B, n))
    sh = 0.2 * rng.standard_normal((B, m))
    l, u = l0[None, :] + sh, u0[None, :] + sh
    t = run(P, A, l0, u0, q, l, u, 4, reps)
    g = run(P, A, l0, u0, q, l, u, 1, reps)
    compare(f"random n={n} m={m} B={B}", t, g)

def mpc_case(N, B, reps=1, big=False):
    cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
    mats = oracle.mpc_build(**{**cfg, "N": N})
    X, U, ref = c2_batch(B, seed=3)
    f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
    t = run(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], f, None, None, 4, reps) if False else None
    B_ = B
    def go(kernel):
        s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=B_, kernel=kernel, **EPS)
        s.update_gradient(f); s.update_upper_bound(ub); s.set_cold_solves(True)
        s.solve(); s.sync(); s.enable_timing(True); s.kernel_ms(reset=True)
        for _ in range(reps): s.solve()
        s.sync(); ms, cnt = s.kernel_ms()
        x, y = s.solution(); info = s.info(); name = s.kernel_name; s.close()
        return x, y, info, ms / max(cnt, 1), name
    t = go(4); g = go(1)
    compare(f"mpc N={N} B={B}", t, g)
    if N == 15:
        compare(f"mpc N={N} B={B} small-vs-generic", go(2), g)
    it = g[2]["iter"].astype(np.int64).sum()
    n, m = N, 2 * N
    for nm, r in (("tile", t), ("generic", g)):
        fl = it * 2.0 * (n * n + 2 * (m // 2) * n)
        print(f"   {nm}: {B / (r[3] * 1e-3):.3e} solves/s, {it / (r[3] * 1e-3):.3e} inst-iter/s, executed {fl / (r[3] * 1e-3) / 1e12:.2f} TFLOP/s", flush=True)

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"):
        mpc_case(15, 512)
        shape_case(12, 20, 64, 1)
        shape_case(17, 33, 40, 2)
        shape_case(40, 70, 37, 3)
        shape_case(100, 200, 64, 4)
        shape_case(200, 400, 24, 5)
        mpc_case(30, 256)
        mpc_case(100, 64)
    if what in ("all", "perf"):
        mpc_case(100, 16384, reps=2)
        shape_case(200, 400, 4096, 6, reps=1)
    if what == "tileperf":
        for N, B in ((100, 16384), (100, 65536), (30, 65536), (50, 65536)):
            cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
            mats = oracle.mpc_build(**{**cfg, "N": N})
            X, U, ref = c2_batch(B, seed=3)
            f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
            s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=B, kernel=4, **EPS)
            s.update_gradient(f); s.update_upper_bound(ub); s.set_cold_solves(True)
            s.solve(); s.sync(); s.enable_timing(True); s.kernel_ms(reset=True)
            for _ in range(2): s.solve()
            s.sync(); ms, cnt = s.kernel_ms(); ms /= cnt
            info = s.info(); it = info["iter"].astype(np.int64).sum(); n, m = N, 2 * N
            print(f"tile N={N} B={B}: {ms:.3f} ms, {B / (ms * 1e-3):.3e} solves/s, iters mean {info['iter'].mean():.1f} max {info['iter'].max()}, "
                  f"executed {it * 2.0 * (n * n + 2 * (m // 2) * n) / (ms * 1e-3) / 1e12:.2f} TFLOP/s, solved {(info['status'] == 1).mean():.4f}", flush=True)
            s.close()
    if what == "c5":
        import time
        N, B, steps = 100, int(sys.argv[2]) if len(sys.argv) > 2 else 65536, int(sys.argv[3]) if len(sys.argv) > 3 else 200
        cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
        conf = dict(Ad=cfg["Ad"], Bd=cfg["Bd"], Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=N)
        X0, U0, _ = c2_batch(B, seed=31); X0 *= 0.2; U0 *= 0.1
        phase = np.random.default_rng(5).integers(0, 200, B).astype(np.int32)
        for kernel in (4,):
            mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=B, kernel=kernel, **EPS)
            mpc.set_state(X=X0, U=U0, ref=np.zeros(B))
            mpc.closed_loop(3, 0.1, 200, phase)     # warm-up
            mpc.set_state(X=X0, U=U0, ref=np.zeros(B)); mpc.solver.reset()
            t0 = time.perf_counter()
            bad, it = mpc.closed_loop(steps, 0.1, 200, phase)
            dt = time.perf_counter() - t0
            n, m = N, 2 * N
            print(f"c5 closed loop N={N} B={B} steps={steps} kernel={mpc.solver.kernel_name}: {dt:.3f} s, {B * steps / dt:.3e} solves/s, "
                  f"not solved {bad}, mean iters {it / (B * steps):.2f}, executed {it * 2.0 * (n * n + 2 * (m // 2) * n) / dt / 1e12:.2f} TFLOP/s", flush=True)
            mpc.close()
    if what == "ncu":
        N, B = 100, 16384
        cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
        mats = oracle.mpc_build(**{**cfg, "N": N})
        X, U, ref = c2_batch(B, seed=3)
        f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
        s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=B, kernel=4, **EPS)
        s.update_gradient(f); s.update_upper_bound(ub); s.set_cold_solves(True)
        s.solve(); s.sync()
        print("ncu run done", s.info()["iter"].mean())
        s.close()
