#!/bin/bash
# Builds oracle/_ref/ref_mpc_gpu: the reference's unmodified ModelPredictiveControlAPI.cpp (from /root/reference,
# never copied) + the PRODUCT OsqpEigen shim + libsolvempc_b200.so.  Only possible in the build container; the
# binary travels to the GPU box with the snapshot (oracle/_ref is git-ignored, not gpurun-ignored).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"; REPO="$(dirname "$HERE")"; REF=/root/reference
OUT="$HERE/_ref/ref_mpc_gpu"
if [ "$OUT" -nt "$HERE/ref_gpu_driver.cpp" ] && [ "$OUT" -nt "$REPO/include/OsqpEigen/OsqpEigen.h" ] && [ "$OUT" -nt "$REPO/include/solvempc_b200.h" ]; then exit 0; fi
mkdir -p "$HERE/_ref"
g++ -O1 -std=c++11 -D_GLIBCXX_USE_CXX11_ABI=0 -DEIGEN_STACK_ALLOCATION_LIMIT=0 -w \
    -I"$REPO/include" -I"$HERE/ref_stubs" -I"$REF/include" \
    -o "$OUT" "$HERE/ref_gpu_driver.cpp" "$REF/src/ModelPredictiveControlAPI.cpp" \
    -L"$REPO/solvempc_b200" -lsolvempc_b200 -Wl,-rpath,'$ORIGIN/../../solvempc_b200'
