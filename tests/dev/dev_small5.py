"""Development driver (GPU): kernel 5 (register-operator DMMA tile) against kernel 2 (one warp per QP) on config 2."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import solvempc_b200 as sm
from problems import c2_batch

EPS = dict(eps_abs=1e-5, eps_rel=1e-5)
cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
mats = oracle.mpc_build(**cfg)


def run(f, ub, kernel, reps=5, sched=True):
    B = f.shape[0]
    s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=B, kernel=kernel, **EPS)
    s.update_gradient(f); s.update_upper_bound(ub); s.set_cold_solves(True); s.set_scheduling(sched)
    s.solve(); s.sync(); s.enable_timing(True); s.kernel_ms(reset=True)
    for _ in range(reps): s.solve()
    s.sync(); ms, cnt = s.kernel_ms(); info = s.info(); x, y = s.solution(); name = s.kernel_name; s.close()
    return ms / cnt, info, x, y, name


for B in (4096, 65536):
    X, U, ref = c2_batch(B, seed=0)
    f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
    res = {}
    for kernel, sched in ((2, True), (5, True), (5, False)):
        ms, info, x, y, name = run(f, ub, kernel, sched=sched)
        res[(kernel, sched)] = (info, x, y)
        it = info["iter"].astype(np.int64)
        print(f"B={B} kernel {kernel} ({name}) sched={sched}: {ms * 1e3:.1f} us, {B / (ms * 1e-3):.3e} solves/s, {it.sum() / (ms * 1e-3):.3e} inst-iter/s, "
              f"iters mean {it.mean():.1f} max {it.max()}, solved {(info['status'] == 1).mean():.4f}", flush=True)
    a, b = res[(2, True)], res[(5, True)]
    sc = np.maximum(np.abs(a[1]).max(axis=1), 1e-9)
    print(f"   5 vs 2: status eq {(a[0]['status'] == b[0]['status']).mean():.4f} iter eq {(a[0]['iter'] == b[0]['iter']).mean():.4f} "
          f"x relerr {(np.abs(a[1] - b[1]).max(axis=1) / sc).max():.2e} y abserr {np.abs(a[2] - b[2]).max():.2e}", flush=True)
    c = res[(5, False)]
    print(f"   5 sched vs index order: x equal {np.array_equal(b[1], c[1])} iter equal {np.array_equal(b[0]['iter'], c[0]['iter'])}", flush=True)

X, U, ref = c2_batch(4096, seed=0)
f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
_, info, _, _, _ = run(f, ub, 2)
hard = int(np.argmax(info["iter"]))
for B in (8, 148 * 8, 148 * 16, 148 * 32, 148 * 64, 148 * 256):
    fb, ubb = np.repeat(f[hard:hard + 1], B, 0), np.repeat(ub[hard:hard + 1], B, 0)
    for kernel in (2, 5):
        ms, info, _, _, _ = run(fb, ubb, kernel, sched=False)
        n_it = int(info["iter"][0])
        print(f"hard x{B} kernel {kernel}: {ms * 1e3:.1f} us, iters {n_it}, {B * n_it / (ms * 1e-3):.3e} inst-iter/s", flush=True)
