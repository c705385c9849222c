# GPU tests against a library built with -DSMPC_DEBUG_BOUNDS (in-kernel checks of the queue / ticket protocols and of the indices
# taken from them; device_types.cuh).  Build first, in the build container, OUT OF TREE (the in-tree .so stays the release build):
#   rm -rf /tmp/dbg && mkdir -p /tmp/dbg/solvempc_b200 /tmp/dbg/include && cp -r solvempc_b200/csrc /tmp/dbg/solvempc_b200/ &&
#   cp -r include/* /tmp/dbg/include/ && rm -rf /tmp/dbg/solvempc_b200/csrc/build &&
#   make -C /tmp/dbg/solvempc_b200/csrc -j8 NVEXTRA=-DSMPC_DEBUG_BOUNDS && cp /tmp/dbg/solvempc_b200/libsolvempc_b200.so build/libsolvempc_b200_debug.so
# then, through gpurun:  bash tools/run_debug_bounds.sh
set -x
export SOLVEMPC_B200_LIB=$PWD/build/libsolvempc_b200_debug.so
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_debug_bounds.log 2>&1
tail -3 gpurun_out/pytest_debug_bounds.log
SMPC_SMALL_FUSED=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_closed_loop.py -m gpu -x -q > gpurun_out/pytest_debug_bounds_fused.log 2>&1
tail -3 gpurun_out/pytest_debug_bounds_fused.log
echo "failed checks printed: $(cat gpurun_out/pytest_debug_bounds.log gpurun_out/pytest_debug_bounds_fused.log | grep -c SMPC_DEBUG_BOUNDS)"
