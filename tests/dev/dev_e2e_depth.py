import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import bench, solvempc_b200 as sm
wl = bench.C2(4096, 0)
wl.setup(sm, torch, 0, 0)
st = torch.cuda.Stream(); wl.mpc.set_stream(st.cuda_stream)
wl.step_e2e(); torch.cuda.synchronize()
for depth in (1, 2, 3, 4, 6):
    ms = [wl.pipelined_e2e(sm, torch, 0, 0, 200, 5, depth=depth) for _ in range(3)]
    print(depth, ["%.1f us" % (1e3 * m) for m in ms], "%.1f M/s" % (4096 / min(ms) / 1e3), flush=True)
