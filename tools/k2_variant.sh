#!/bin/bash
# Experiment build of the single-large-permanent kernels only: tools/k2_variant.sh TAG "<extra nvcc flags>"
# -> lightworks_b200/variants/liblwb200_TAG.so (other units are linked from the base build).
set -e
TAG=$1; shift
cd "$(dirname "$0")/../lightworks_b200/csrc"
mkdir -p build_$TAG ../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
for u in k2_g0 k2_g1 k2_g2 k2_g3; do
  nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c $u.cu -o build_$TAG/$u.o &
done
wait
OBJS=$(ls build/*.o | grep -v -E '/k2_g[0-9]\.o$')
nvcc $ARCH -shared -o ../variants/liblwb200_$TAG.so $OBJS build_$TAG/*.o -ldl
echo built $TAG
