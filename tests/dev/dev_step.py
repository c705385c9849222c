"""Development driver (GPU): config-2 controller step time (set_state from device buffers + controllerStep) with and without
the library's kernel-timing events (which sit between the kernels and disable the programmatic dependent launch overlap)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import solvempc_b200 as sm
from problems import c2_batch

B = 4096
X, U, ref = c2_batch(B, seed=0)
mpc = sm.BatchedModelPredictiveControlAPI(os.path.join(ROOT, "config", "MPC_API.json"), batch=B, eps_abs=1e-5, eps_rel=1e-5)
mpc.set_stream(torch.cuda.current_stream().cuda_stream)
mpc.solver.set_cold_solves(True)
d = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (X, U, ref)]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
for timing in (False, True, False):
    mpc.solver.enable_timing(timing)
    for _ in range(5):
        mpc.set_state(X=d[0], U=d[1], ref=d[2]); mpc.controller_step_async()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
    for e0, e1 in evs:
        flush.zero_()
        e0.record()
        mpc.set_state(X=d[0], U=d[1], ref=d[2]); mpc.controller_step_async()
        e1.record()
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in evs])
    k = mpc.solver.kernel_ms() if timing else (0.0, 1)
    print(f"timing events {'on ' if timing else 'off'}: step {1e3 * ms.mean():.1f} us (min {1e3 * ms.min():.1f}), kernel {1e3 * k[0] / max(k[1], 1):.1f} us, solved {mpc.solver.count_solved()}", flush=True)
