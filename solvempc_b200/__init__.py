"""solvempc_b200 -- B200-native (sm_100a) batched MPC QP solver behind solveMPC's ModelPredictiveControlAPI.

The product is libsolvempc_b200.so (hand-written CUDA + a C ABI, include/solvempc_b200.h); this package is the
thin Python host mirror used by the tests and bench.py.  There is no CPU fallback.
"""
from ._lib import (DEVICE, HOST, LIB_PATH, SOLVED, SOLVED_INACCURATE, MAX_ITER_REACHED, PRIMAL_INFEASIBLE,
                   DUAL_INFEASIBLE, UNSOLVED, PRIMAL_INFEASIBLE_INACCURATE, DUAL_INFEASIBLE_INACCURATE,
                   Settings, SolveMpcError, default_settings, lib)
from .sharding import gather_results, shard_arrays, shard_bounds
from .solver import BatchedMimoMPC, BatchedModelPredictiveControlAPI, BatchedSolver, shared_plan_inspect

__all__ = ["BatchedMimoMPC", "BatchedModelPredictiveControlAPI", "BatchedSolver", "shared_plan_inspect", "default_settings",
           "Settings", "SolveMpcError", "lib", "HOST", "DEVICE", "LIB_PATH", "shard_bounds", "shard_arrays", "gather_results"]
