# Development: SMPC_TILE_PROFILE build (tools/ab_tile.sh prof "-DSMPC_TILE_PROFILE"): cycle split of tile 0 on configs 3 and 5
export SOLVEMPC_B200_LIB=$PWD/build/ab/lib_prof.so
python tests/dev/dev_tile.py c5 65536 2 > gpurun_out/prof_c5.txt 2>&1
grep -n "tile profile" gpurun_out/prof_c5.txt | tail -2
tail -45 gpurun_out/prof_c5.txt
