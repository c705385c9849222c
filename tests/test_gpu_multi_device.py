"""ONE global batch split over several solver handles / devices gives bitwise the single-handle result.

On a multi-GPU box the shards go to different devices (one handle per device from one process: the C ABI takes `device`);
on a one-GPU box the same code runs with several handles on device 0, which exercises the same partition / gather logic.
The C++ statement of the same thing (examples/multi_gpu_batch.cpp, built by __graft_entry__.build()) is run as well."""
import json
import os
import subprocess

import numpy as np
import pytest

import solvempc_b200 as sm
from problems import c2_batch
from solvempc_b200.sharding import shard_bounds

pytestmark = pytest.mark.gpu
EPS = dict(eps_abs=1e-5, eps_rel=1e-5)


def _step(cfg_path, device, X, U, ref, kernel=0):
    mpc = sm.BatchedModelPredictiveControlAPI(cfg_path, batch=X.shape[0], device=device, kernel=kernel, **EPS)
    mpc.set_state(X=X, U=U, ref=ref)
    mpc.controller_step_async()
    return mpc


@pytest.mark.parametrize("kernel", [2, 4, 1])
def test_one_global_batch_over_all_devices_is_bitwise_the_single_device_run(repo_root, kernel):
    cfg_path = os.path.join(repo_root, "config", "MPC_API.json")
    devices = sm.lib().smpc_device_count()
    world = devices if devices > 1 else 3
    B = 1021 if kernel != 1 else 203                                  # ragged over any world size
    X, U, ref = c2_batch(B, seed=77)
    handles = []
    for r in range(world):                                            # all shards in flight at once, one handle per shard
        lo, hi = shard_bounds(B, world, r)
        handles.append((lo, hi, _step(cfg_path, r % devices, np.ascontiguousarray(X[lo:hi]), U[lo:hi].copy(), ref[lo:hi].copy(), kernel)))
    Ug, xg, stg, itg = np.empty(B), np.empty((B, 15)), np.empty(B, np.int32), np.empty(B, np.int32)
    for lo, hi, mpc in handles:                                       # the "gather": every shard fills its rows
        _, Ug[lo:hi] = mpc.state()
        xg[lo:hi] = mpc.solver.solution(want_y=False)
        info = mpc.solver.info()
        stg[lo:hi], itg[lo:hi] = info["status"], info["iter"]
        mpc.close()
    one = _step(cfg_path, 0, X, U, ref, kernel)
    _, U1 = one.state()
    x1 = one.solver.solution(want_y=False)
    info = one.solver.info()
    one.close()
    assert (stg == 1).all()
    assert np.array_equal(Ug, U1) and np.array_equal(xg, x1) and np.array_equal(stg, info["status"]) and np.array_equal(itg, info["iter"])


def test_generic_kernel_large_smem_on_every_device(ref_mats):
    """the > 48 KB dynamic shared memory attribute of the generic kernel is per device; a solver on device 1 created
    after one on device 0 must still launch (horizon 100: n = 100, m = 200 needs ~ 130 KB per CTA)."""
    import oracle
    _, cfg = ref_mats
    devices = sm.lib().smpc_device_count()
    conf = dict(Ad=cfg["Ad"], Bd=cfg["Bd"], Cd=cfg["Cd"], K=cfg["K"], Q=cfg["Q"], R=cfg["R"], RD=cfg["RD"], horizon=100)
    X, U, ref = c2_batch(4, seed=3)
    outs = []
    for dev in range(min(devices, 2)) if devices > 1 else (0, 0):
        mpc = sm.BatchedModelPredictiveControlAPI(conf, batch=4, device=dev, kernel=1, **EPS)
        mpc.set_state(X=X * 0.2, U=U * 0.1, ref=np.zeros(4))
        assert mpc.controllerStep()
        outs.append(mpc.solver.solution(want_y=False))
        mpc.close()
    assert np.array_equal(outs[0], outs[1])


def test_cpp_multi_device_example(repo_root):
    exe = os.path.join(repo_root, "examples", "_build", "multi_gpu_batch")
    if not os.path.exists(exe):
        pytest.skip("examples/_build/multi_gpu_batch was not built (run __graft_entry__.build())")
    devices = sm.lib().smpc_device_count()
    out = subprocess.run([exe, os.path.join(repo_root, "config", "MPC_API.json"), "4099", str(devices if devices > 1 else 3)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    res = json.loads(out.stdout)
    assert res["bitwise_equal"] and res["solved"] == 4099
