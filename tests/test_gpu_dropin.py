"""Drop-in proof (BASELINE config 1): the reference's OWN unmodified ModelPredictiveControlAPI.cpp, compiled in the
build container against the product shim include/OsqpEigen/OsqpEigen.h (oracle/build_ref_gpu.sh), runs its
controllerStep on the GPU through libsolvempc_b200.so and reproduces the golden run."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_reference_class_on_the_gpu_solver(repo_root, golden):
    exe = os.path.join(repo_root, "oracle", "_ref", "ref_mpc_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_mpc_gpu was not built (needs /root/reference at build time)")
    env = dict(os.environ, SOLVEMPC_EPS="1e-5")
    out = subprocess.run([exe], cwd=repo_root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = json.loads(out.stdout)
    for g, c in zip(got["cases"], golden["cases"]):
        assert g["ok"] and g["status"] == c["status"] == 1 and g["iter"] == c["iter"]
        x, xo = np.array(g["x"]), np.array(c["x"])
        assert np.abs(x - xo).max() <= 1e-7 * np.abs(xo).max()
        assert abs(g["U_after"] - c["U_after"]) < 1e-9
    cl = golden["closed_loop"]
    assert got["closed_loop"]["iters"] == cl["iters"]
    assert np.abs(np.array(got["closed_loop"]["U"]) - np.array(cl["U"])).max() < 1e-8
