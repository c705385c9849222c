"""Small invocation of every kernel family for compute-sanitizer (memcheck / racecheck) runs:
compute-sanitizer --tool memcheck python tests/dev/sanitize.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import solvempc_b200 as sm
from problems import c2_batch, c3_batch, c4_plants, random_qp

EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
cfgp = os.path.join(ROOT, "config", "MPC_API.json")
cfg = oracle.load_config(cfgp)
for kernel in (2, 5, 1, 4):                                      # shared-factor kernels through the MPC layer
    B = 37
    X, U, ref = c2_batch(B, seed=1)
    mpc = sm.BatchedModelPredictiveControlAPI(cfgp, batch=B, kernel=kernel, **EPS)
    mpc.set_state(X=X, U=U, ref=ref)
    ok = mpc.controllerStep(); mpc.controllerStep()
    print(mpc.solver.kernel_name, "solved", mpc.solver.count_solved(), flush=True)
    mpc.close()
conf = dict(Ad=cfg["Ad"], Bd=cfg["Bd"], Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=40)
mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=19, **EPS)   # tile kernel, closed loop with a graph
X, U, ref = c2_batch(19, seed=2)
mpc.set_state(X=X * 0.2, U=U * 0.1, ref=np.zeros(19))
print("closed loop", mpc.closed_loop(3, 0.1, 6, np.arange(19, dtype=np.int32) % 6), mpc.solver.kernel_name, flush=True)
mpc.close()
P, q, A, l, u = random_qp(20, 35, seed=3)                        # unpaired tile kernel, infeasible + equality rows
l[:5] = u[:5]; l[7], u[7] = -np.inf, np.inf
s = sm.BatchedSolver(P, A, l, u, batch=11, kernel=4, **EPS)
s.update_gradient(q[None, :] + 0.2 * np.random.default_rng(0).standard_normal((11, 20))); s.solve()
print(s.kernel_name, s.info()["status"], flush=True); s.close()
x0, xr = c3_batch(9, seed=0)                                     # multi-input layer
mm = sm.BatchedMimoMPC(os.path.join(ROOT, "config", "quadrotor.json"), batch=9, **EPS)
mm.set_state(x0=x0, xr=xr); print("mimo solved", mm.controllerStep(), flush=True); mm.close()
Ad, Bd = c4_plants(13, cfg, seed=2)                              # per-instance regime, paired and unpaired
conf = dict(Ad=Ad, Bd=Bd, Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=30, per_instance=1)
mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=13, **EPS)
X, U, ref = c2_batch(13, seed=31)
mpc.set_state(X=X, U=U, ref=ref); print(mpc.solver.kernel_name, "pairs", mpc.solver.row_pairs, mpc.controllerStep(), flush=True); mpc.close()
Ps = np.array([random_qp(10, 16, seed=50 + b)[0] for b in range(7)]); As = np.array([random_qp(10, 16, seed=50 + b)[2] for b in range(7)])
s = sm.BatchedSolver.batched(Ps, As, np.full(16, -2.0), np.full(16, 2.0), **EPS)
s.update_gradient(np.random.default_rng(1).standard_normal((7, 10))); s.solve(); print("instance unpaired", s.info()["status"], flush=True); s.close()
Ps = np.array([random_qp(40, 50, seed=70 + b)[0] for b in range(3)]); As = np.array([random_qp(40, 50, seed=70 + b)[2] for b in range(3)])
s = sm.BatchedSolver.batched(Ps, As, np.full(50, -2.0), np.full(50, 2.0), **EPS)   # generic per-instance kernel (n > 32)
s.update_gradient(np.random.default_rng(2).standard_normal((3, 40))); s.solve(); print("instance generic", s.info()["status"], flush=True); s.close()
print("sanitize run complete")
