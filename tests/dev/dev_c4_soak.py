"""Development driver (GPU): the per-instance pair kernel against the CPU oracle's per-plant controllers (oracle/batch_drivers.c) on
a few thousand distinct plants: statuses and iteration counts must be identical, solutions equal to round-off."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import solvempc_b200 as sm
from problems import c2_batch, c4_plants

EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
for N, B, seed in ((30, 8192, 5), (15, 4096, 6), (12, 2048, 7)):
    Ad, Bd = c4_plants(B, cfg, seed=seed)
    X, U, ref = c2_batch(B, seed=seed + 100)
    conf = dict(Ad=Ad, Bd=Bd, Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=N, per_instance=1)
    mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=B, **EPS)
    mpc.set_state(X=X, U=U, ref=ref)
    mpc.controller_step_async()
    x, _ = mpc.solver.solution(); info = mpc.solver.info(); name = mpc.solver.kernel_name
    mpc.close()
    out = oracle.plant_batch(cfg, Ad, Bd, X, U, ref, N, settings=oracle.default_settings(**EPS), nthreads=os.cpu_count() or 1)
    st_eq = (info["status"] == out["status"]).mean(); it_eq = (info["iter"] == out["iter"]).mean()
    ok = out["status"] == 1
    err = (np.abs(x[ok] - out["x"][ok]).max(axis=1) / np.maximum(np.abs(out["x"][ok]).max(axis=1), 1e-9)).max()
    print(f"N={N} B={B} {name}: status equal {st_eq:.5f}, iterations equal {it_eq:.5f}, max rel err on SOLVED {err:.2e}, "
          f"iters mean {info['iter'].mean():.1f}, rho updates mean {info['rho_updates'].mean():.2f}", flush=True)
