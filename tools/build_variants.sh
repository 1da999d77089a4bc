#!/bin/bash
# dev tool: A/B builds of ONE translation unit with -D knobs, linked against the regular objects.
#   tools/build_variants.sh kf_p1 name1 "-DKFB_P1_SLOTS=8" name2 "-DKFB_P1_MINB=10" ...
# -> build/variants/libkfb200_<name>.so   (select with KFB_LIB=... ; pymc_statespace_b200/_lib.py)
set -e
cd "$(dirname "$0")/.."
TU=$1; shift
rm -rf build/variants; mkdir -p build/variants
OTHERS=$(ls build/csrc/*.o | grep -v "/${TU}.o")
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
      -Xcompiler -fvisibility=hidden -Xptxas -v $flags -c pymc_statespace_b200/csrc/${TU}.cu -o build/variants/${TU}_${name}.o 2> build/variants/${TU}_${name}.ptxas.log &
done
wait
for o in build/variants/${TU}_*.o; do
  name=$(basename $o .o); name=${name#${TU}_}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libkfb200_${name}.so $OTHERS $o
done
