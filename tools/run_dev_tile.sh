set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python tools/dev_tile.py parity > gpurun_out/dev_tile_parity.log 2>&1; echo "parity rc=$?"
tail -30 gpurun_out/dev_tile_parity.log
timeout 600 python tools/dev_tile.py perf > gpurun_out/dev_tile_perf.log 2>&1; echo "perf rc=$?"
tail -12 gpurun_out/dev_tile_perf.log
