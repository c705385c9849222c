# Development: SMPC_TILE_PROFILE builds (tools/ab_tile.sh <name> "-DSMPC_TILE_PROFILE ..."): cycle split of tile 0 on config 3
for v in "$@"; do
export SOLVEMPC_B200_LIB=$PWD/build/ab/lib_$v.so
echo "== $v"
timeout 300 python tests/dev/dev_c3.py 32768 > gpurun_out/prof_$v.txt 2>&1
grep "tile profile" gpurun_out/prof_$v.txt | tail -1
grep "warp .* iteration" gpurun_out/prof_$v.txt | tail -2
done
