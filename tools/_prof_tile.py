import os, sys
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import solvempc_b200 as sm
from problems import c2_batch, c3_batch
EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
N = 100
mats = oracle.mpc_build(**{**cfg, "N": N})
X, U, ref = c2_batch(2368, seed=3)
f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=2368, kernel=4, **EPS)
s.update_gradient(f); s.update_upper_bound(ub); s.set_cold_solves(True)
s.solve(); s.sync(); print("N=100 iters mean", s.info()["iter"].mean(), flush=True); s.close()
B = 1184
x0, xr = c3_batch(B, seed=1)
mpc = sm.BatchedMimoMPC(os.path.join(ROOT, "config", "quadrotor.json"), batch=B, **EPS)
mpc.set_state(x0=x0, xr=xr); mpc.controller_step_async(); mpc.solver.sync()
print("c3 iters mean", mpc.solver.info()["iter"].mean(), flush=True)
