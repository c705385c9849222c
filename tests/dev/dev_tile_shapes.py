import os, sys
ROOT="/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracle, solvempc_b200 as sm
from problems import c2_batch
EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
for N, B in ((30, 65536), (50, 65536), (64, 32768), (100, 32768)):
    mats = oracle.mpc_build(**{**cfg, "N": N})
    X, U, ref = c2_batch(B, seed=3)
    f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
    s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=B, kernel=4, **EPS)
    s.update_gradient(f); s.update_upper_bound(ub); s.set_cold_solves(True)
    s.solve(); s.sync(); s.enable_timing(True); s.kernel_ms(reset=True)
    for _ in range(3): s.solve()
    s.sync(); ms, cnt = s.kernel_ms(); info = s.info(); s.close()
    print(f"N={N} B={B}: {ms/cnt:.3f} ms  iters mean {info['iter'].mean():.1f} solved {(info['status']==1).mean():.4f} checksum {int(info['iter'].sum())}", flush=True)
