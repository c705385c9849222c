"""Development driver (GPU): split of the small-QP kernel's time per iteration into the iteration chain and the
termination checks, unloaded (B = 1) and at 12 warps per SM (B = 1776): the same instance replicated B times, run for a
fixed number of iterations with and without the periodic check / rho adaptation."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import solvempc_b200 as sm
from problems import c2_batch

cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
mats = oracle.mpc_build(**cfg)


def timed(f, ub, reps=5, **kw):
    B = f.shape[0]
    s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=B, kernel=2, **kw)
    s.update_gradient(f); s.update_upper_bound(ub); s.set_cold_solves(True); s.set_scheduling(False)
    s.solve(); s.sync(); s.enable_timing(True); s.kernel_ms(reset=True)
    for _ in range(reps): s.solve()
    s.sync(); ms, cnt = s.kernel_ms(); info = s.info(); s.close()
    return ms / cnt, info


X, U, ref = c2_batch(64, seed=0)
f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
for B in (1, 1776):
    fb, ubb = np.repeat(f[:1], B, 0), np.repeat(ub[:1], B, 0)
    res = {}
    for tag, kw in (("no checks 200", dict(max_iter=200, check_termination=0, adaptive_rho=0, eps_abs=1e-30, eps_rel=1e-30)),
                    ("no checks 400", dict(max_iter=400, check_termination=0, adaptive_rho=0, eps_abs=1e-30, eps_rel=1e-30)),
                    ("checks/25 200", dict(max_iter=200, check_termination=25, adaptive_rho=0, eps_abs=1e-30, eps_rel=1e-30)),
                    ("checks/25 400", dict(max_iter=400, check_termination=25, adaptive_rho=0, eps_abs=1e-30, eps_rel=1e-30)),
                    ("checks+adapt/25 200", dict(max_iter=200, check_termination=25, adaptive_rho=1, adaptive_rho_interval=25, eps_abs=1e-30, eps_rel=1e-30)),
                    ("checks+adapt/25 400", dict(max_iter=400, check_termination=25, adaptive_rho=1, adaptive_rho_interval=25, eps_abs=1e-30, eps_rel=1e-30))):
        ms, info = timed(fb, ubb, **kw)
        res[tag] = ms
        print(f"B={B} {tag}: {ms * 1e3:.1f} us  iters {int(info['iter'][0])}", flush=True)
    for k in ("no checks", "checks/25", "checks+adapt/25"):
        d = (res[k + " 400"] - res[k + " 200"]) * 1e-3 / 200 * 1.965e9
        print(f"B={B} {k}: {d:.0f} cycles per iteration", flush=True)
