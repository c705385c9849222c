# Development: builds A/B variants of the library that differ in the tile kernel's compile-time switches only
# (build/ab/lib_<name>.so; pick one with SOLVEMPC_B200_LIB).  usage: bash tools/ab_tile.sh name "-DFLAG ..." [name "-D..."]...
set -e
cd "$(dirname "$0")/.."
make -C solvempc_b200/csrc -s
mkdir -p build/ab
C=solvempc_b200/csrc
OTHERS=$(ls $C/build/*.o | grep -v admm_shared_tile.o)
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -diag-suppress 128 -Xcompiler -fPIC --fmad=true $flags -c -o build/ab/tile_$name.o $C/admm_shared_tile.cu
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/lib_$name.so build/ab/tile_$name.o $OTHERS -lcudart_static -lpthread -ldl -lrt
  echo built build/ab/lib_$name.so
done
