# Development: times config 3 (cold quadrotor batch) and config 5 (closed loop) with each A/B library of tools/ab_tile.sh
for v in "$@"; do
  if [ "$v" = "tree" ]; then unset SOLVEMPC_B200_LIB; else export SOLVEMPC_B200_LIB=$PWD/build/ab/lib_$v.so; fi
  echo "== $v"
  python tests/dev/dev_c3.py 131072 2>&1 | tail -1
  python tests/dev/dev_tile.py c5 65536 100 2>&1 | tail -1
done
