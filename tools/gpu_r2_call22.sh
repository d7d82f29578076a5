#!/bin/bash
# K3r: sensitivity to the order in which CTAs take items (L2 locality probe)
mkdir -p gpurun_out
{
echo "== base (lex order)"; C3_ONLY=1 python tools/slos_profile.py 5 1 4
for v in ord1 ord2; do echo "== variant $v"; C3_ONLY=1 LWB_LIB_PATH=lightworks_b200/variants/liblwb200_$v.so python tools/slos_profile.py 5 1 4; done
} > gpurun_out/c22_slos.log 2>&1
cat gpurun_out/c22_slos.log
