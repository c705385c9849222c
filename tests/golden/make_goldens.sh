#!/bin/bash
# Regenerates tests/golden/assembly_ref.json from the REFERENCE'S OWN sources
# (/root/reference/src/ModelPredictiveControlAPI.cpp compiled where it lies against the
# recording stubs in oracle/ref_stubs; SURVEY.md 8c recipe).  Only runnable in the build
# container (needs /root/reference); the committed JSON is what travels to the GPU box.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/../.." && pwd)"
make -C "$REPO/oracle" -s ref
RUN="$(mktemp -d)"
ln -s /root/reference/config "$RUN/config"     # the shipped config/MPC_API.json, read cwd-relative (cpp:12)
( cd "$RUN" && ORC_EPS=1e-5 ORC_RHO_INTERVAL=25 "$REPO/oracle/_ref/ref_dump" ) > "$REPO/tests/golden/assembly_ref.json"
rm -rf "$RUN"
"$REPO/oracle/_ref/ref_wire" > "$REPO/tests/golden/wire_ref.json"   # reference SerialPort parsing / formatting
python -c "import json;d=json.load(open('$REPO/tests/golden/assembly_ref.json'));print('golden ok: n=%d m=%d cases=%d'%(d['n'],d['m'],len(d['cases'])))"
