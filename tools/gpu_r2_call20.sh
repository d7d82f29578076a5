#!/bin/bash
# K3r: cache policy of the prefix-stream loads
mkdir -p gpurun_out
{
echo "== base (plain ld.global)"; python tools/slos_profile.py 5 1 4
for v in ld1 ld2 ld3; do echo "== variant $v"; LWB_LIB_PATH=lightworks_b200/variants/liblwb200_$v.so python tools/slos_profile.py 5 1 4; done
echo "== base again"; python tools/slos_profile.py 5 1 4
} > gpurun_out/c20_slos.log 2>&1
cat gpurun_out/c20_slos.log
