"""Development driver (GPU): config 3 (quadrotor N=50, n=200, m=400) throughput of the shared-factor kernels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import solvempc_b200 as sm
from problems import c3_batch

EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
path = os.path.join(ROOT, "config", "quadrotor.json")
for B in [int(a) for a in sys.argv[1:]] or [8192]:
    x0, xr = c3_batch(B, seed=1)
    mpc = sm.BatchedMimoMPC(path, batch=B, **EPS)
    mpc.solver.set_cold_solves(True)
    mpc.set_state(x0=x0, xr=xr)
    mpc.controller_step_async(); mpc.solver.sync()
    mpc.solver.enable_timing(True); mpc.solver.kernel_ms(reset=True)
    for _ in range(2): mpc.controller_step_async()
    mpc.solver.sync()
    ms, cnt = mpc.solver.kernel_ms(); ms /= cnt
    info = mpc.solver.info(); it = info["iter"].astype(np.int64)
    n, m = mpc.n_variables, mpc.n_constraints
    print(f"c3 B={B} {mpc.solver.kernel_name}: {ms:.2f} ms, {B / (ms * 1e-3):.3e} solves/s, iters mean {it.mean():.1f} max {it.max()}, "
          f"solved {(info['status'] == 1).mean():.4f}, executed {it.sum() * 2.0 * (n * n + 2 * (m // 2) * n) / (ms * 1e-3) / 1e12:.2f} TFLOP/s", flush=True)
    mpc.close()
