"""Development driver (GPU): per-iteration latency of the small-QP kernel as a function of resident warps.
Replicates ONE config-2 instance B times (identical iteration counts, no tail) and times the ADMM kernel."""
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


def timed(f, ub, kernel=2, reps=5, sched=False):
    B = f.shape[0]
    s = sm.BatchedSolver(mats["H"], mats["Gbar"], mats["lb"], mats["W0"], batch=B, kernel=kernel, **EPS)
    s.update_gradient(f); s.update_upper_bound(ub); s.set_cold_solves(True); s.set_scheduling(sched)
    s.solve(); s.sync(); s.enable_timing(True); s.kernel_ms(reset=True)
    for _ in range(reps): s.solve()
    s.sync(); ms, cnt = s.kernel_ms(); info = s.info(); s.close()
    return ms / cnt, info


def main():
    X, U, ref = c2_batch(4096, seed=0)
    f, ub = oracle.mpc_batch_vectors(mats, X, U, ref)
    ms, info = timed(f, ub, sched=True)
    it = info["iter"]
    print(f"c2 B=4096 sched: {ms * 1e3:.1f} us, iters mean {it.mean():.1f} max {it.max()}", flush=True)
    ms, info = timed(f, ub, sched=False)
    print(f"c2 B=4096 index order: {ms * 1e3:.1f} us", flush=True)
    hard = int(np.argmax(it)); med = int(np.argsort(it)[len(it) // 2])
    for tag, idx in (("hard", hard), ("median", med)):
        for B in (1, 32, 148 * 4, 148 * 8, 148 * 12, 148 * 24, 148 * 48, 148 * 96):
            fb, ubb = np.repeat(f[idx:idx + 1], B, 0), np.repeat(ub[idx:idx + 1], B, 0)
            ms, info = timed(fb, ubb)
            n_it = int(info["iter"][0])
            waves = max(1, -(-B // (148 * 12)))
            cyc = ms * 1e-3 * 1.965e9 / n_it
            print(f"{tag} iters={n_it} B={B}: {ms * 1e3:.1f} us, {cyc:.0f} cycles per iteration-wave-set, "
                  f"{B * n_it / (ms * 1e-3):.3e} inst-iter/s ({B * n_it * 2250 / (ms * 1e-3) / 1e12:.2f} TFLOP/s executed)", flush=True)


if __name__ == "__main__":
    main()
