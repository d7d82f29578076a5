#!/bin/bash
# K2 variants at 2 warps per scheduler: product chains / tree / static unroll
mkdir -p gpurun_out
{
echo "== base"; python tools/k2_profile.py 22 24 26 28
for v in ch3 ch4 tree u3; do echo "== variant $v"; LWB_LIB_PATH=lightworks_b200/variants/liblwb200_$v.so python tools/k2_profile.py 22 24 26 28; done
} > gpurun_out/c27_k2.log 2>&1
cat gpurun_out/c27_k2.log
