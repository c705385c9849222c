# Profiles kept under profiles/: launch list of the default bench command + one full capture per ADMM kernel at the
# bench's own batch sizes (so that dram bytes per launch = roofline.traffic of that configuration).
# usage (on the GPU box, through gpurun): bash tools/run_ncu.sh [configs, default "c2 c3 c4 c5"]   (gpurun copies back at most
# 64 MiB per call: the four captures together exceed it, so run "c2 c3 c4" and "c5" in two calls)
set -x
CFGS="${*:-c2 c3 c4 c5}"
case " $CFGS " in *" c2 "*) ;; *) SKIP2=1;; esac
A2="--config c2 --steps 3 --warmup 3 --cpu-seconds 0.2"
if [ -z "$SKIP2" ]; then
python bench.py $A2 > gpurun_out/plain_c2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py $A2 > gpurun_out/ncu_launches_c2.log 2>&1
python bench.py $A2 > gpurun_out/plain_c2b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:admm_shared_small -s 3 -c 1 -f -o gpurun_out/prof_c2 python bench.py $A2 > gpurun_out/ncu_c2.log 2>&1
fi
for c in $CFGS; do
  [ $c = c2 ] && continue
  case $c in c3) K=admm_shared_tile;; c4) K=admm_instance;; c5) K=admm_shared_tile;; esac
  A="--config $c --steps 2 --warmup 3 --cpu-seconds 0.2"
  python bench.py $A > gpurun_out/plain_$c.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_$c python bench.py $A > gpurun_out/ncu_$c.log 2>&1
  echo "$c ncu rc=$?"
done
ls -la gpurun_out/*.ncu-rep
