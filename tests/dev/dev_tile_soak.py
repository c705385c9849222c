"""Development driver (GPU): the tile kernel (two tiles per SM where they fit) against the CPU oracle on cold MPC batches."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import solvempc_b200 as sm
from problems import c2_batch

EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
for N, B in ((30, 4096), (50, 4096), (64, 2048), (100, 2048)):
    mats = oracle.mpc_build(**{**cfg, "N": N})
    X, U, ref = c2_batch(B, seed=40 + N)
    f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
    s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=B, kernel=4, **EPS)
    s.update_gradient(f); s.update_upper_bound(ub); s.solve()
    x, _ = s.solution(); info = s.info(); name = s.kernel_name
    s.close()
    ora = oracle.solve_batch(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], f, ub, nthreads=os.cpu_count() or 1, **EPS)
    ok = ora["status"] == 1
    err = (np.abs(x[ok] - ora["x"][ok]).max(axis=1) / np.maximum(np.abs(ora["x"][ok]).max(axis=1), 1e-12)).max()
    print(f"N={N} B={B} {name}: status equal {(info['status'] == ora['status']).mean():.5f}, iterations equal "
          f"{(info['iter'] == ora['iter']).mean():.5f}, max rel err {err:.2e}, iters mean {info['iter'].mean():.1f}", flush=True)
