# Round-end check of the default configuration: GPU tests, smoke, the bench line, its ncu launch list and one full capture of
# the ADMM kernel.  usage (on the GPU box, through gpurun): bash tools/run_ncu_c2.sh
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --config c2 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.err
A2="--config c2 --steps 3 --warmup 3 --cpu-seconds 0.2"
python bench.py $A2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py $A2 > gpurun_out/ncu_launches_c2.log 2>&1
python bench.py $A2 > gpurun_out/plain_c2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:admm_shared_small -s 3 -c 1 -f -o gpurun_out/prof_c2 python bench.py $A2 > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep
