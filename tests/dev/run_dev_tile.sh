set -x
timeout 600 python tests/dev/dev_tile.py parity > gpurun_out/dev_tile_parity.log 2>&1; echo "parity rc=$?"
grep -c "status eq 1.0000" gpurun_out/dev_tile_parity.log; grep "iter eq" gpurun_out/dev_tile_parity.log | awk '{print $1,$2,$3,$4,$8,$9}'
timeout 600 python tests/dev/dev_tile.py tileperf > gpurun_out/dev_tile_perf.log 2>&1; echo "perf rc=$?"
cat gpurun_out/dev_tile_perf.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:admm_shared_tile -c 1 -o gpurun_out/prof_tile_v2 -f python tests/dev/dev_tile.py ncu > gpurun_out/ncu_tile.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_tile.log
