#!/bin/bash
# A/B builds of the fused kernel: libippl_b200_<name>.so with different -D knobs for fused.cu (other objects are shared).
# usage: scripts/build_variants.sh name1 "<flags1>" name2 "<flags2>" ...
set -e
cd "$(dirname "$0")/../ippl_b200/csrc"
make -s
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -ccbin /usr/bin/g++"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  $NV $flags -c fused.cu -o build/fused_$name.o
  objs=$(ls build/*.o | grep -v "build/fused")
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libippl_b200_$name.so $objs build/fused_$name.o -lcufft -lnccl -ccbin /usr/bin/g++
  echo "built libippl_b200_$name.so ($flags)"
done
