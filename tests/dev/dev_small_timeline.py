"""Development driver (GPU, library built with -DSMPC_SMALL_TIMELINE): when does every config-2 instance start and end,
on which SM, and how loaded is the machine over time."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
import solvempc_b200 as sm
from problems import c2_batch

cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
mats = oracle.mpc_build(**cfg)
X, U, ref = c2_batch(4096, seed=0)
f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=4096, kernel=2, eps_abs=1e-5, eps_rel=1e-5)
s.update_gradient(f); s.update_upper_bound(ub); s.set_cold_solves(True); s.set_scheduling(True)
for _ in range(3):
    s.solve(); s.sync()
info = s.info()
t0, t1, place, it = info["pri_res"], info["dua_res"], info["obj"], info["iter"]
base = t0.min()
t0 = (t0 - base) * 1e-3; t1 = (t1 - base) * 1e-3
print(f"span {t1.max():.1f} us; instances {len(it)}; quiet-warp instances {(place % 1 > 0).sum()}")
order = np.argsort(-t1)[:12]
for b in order:
    print(f"  end {t1[b]:6.1f} start {t0[b]:6.1f} dur {t1[b]-t0[b]:6.1f} iters {it[b]:3d} us/iter {(t1[b]-t0[b])/it[b]:.3f} sm {int(place[b])//2048} quiet {place[b] % 1 > 0}")
for lo in range(0, int(t1.max()) + 10, 10):
    act = ((t0 <= lo) & (t1 > lo)).sum()
    print(f"  t={lo:3d} us: {act} instances in flight")
q = place % 1 > 0
if q.any():
    print(f"quiet: iters mean {it[q].mean():.0f} max {it[q].max()}, us/iter mean {((t1-t0)[q]/it[q]).mean():.3f}; end max {t1[q].max():.1f}")
print(f"normal: us/iter mean {((t1-t0)[~q]/it[~q]).mean():.3f}; end max {t1[~q].max():.1f}")
warp = place.astype(np.int64)
per = {}
for b in range(len(it)):
    per.setdefault(int(warp[b]) % 2048, []).append(b)
gaps = []
for v in per.values():
    v = sorted(v, key=lambda b: t0[b])
    gaps += [t0[v[k + 1]] - t1[v[k]] for k in range(len(v) - 1)]
gaps = np.array(gaps)
print(f"gap between the end of an instance and the start of the warp's next: mean {gaps.mean():.2f} us, median {np.median(gaps):.2f}, p90 {np.percentile(gaps, 90):.2f}, max {gaps.max():.2f} ({len(gaps)} gaps)")
first = np.array([min(t0[b] for b in v) for v in per.values()])
print(f"first start per warp: median {np.median(first):.2f} us, max {first.max():.2f}")
tot = np.array([sum(it[v]) for v in per.values()])
print(f"warps used {len(per)}; iterations per warp mean {tot.mean():.0f} max {tot.max()} min {tot.min()}")
s.close()
