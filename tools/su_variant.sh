#!/bin/bash
# Experiment build of the SLOS prefix-unit kernel only: tools/su_variant.sh TAG "<extra nvcc flags>"
# -> lightworks_b200/variants/liblwb200_TAG.so (other units are linked from the base build).
set -e
TAG=$1; shift
cd "$(dirname "$0")/../lightworks_b200/csrc"
mkdir -p build_$TAG ../variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c slos_units.cu -o build_$TAG/slos_units.o
OBJS=$(ls build/*.o | grep -v '/slos_units\.o$')
nvcc $ARCH -shared -o ../variants/liblwb200_$TAG.so $OBJS build_$TAG/slos_units.o -ldl
echo built $TAG
