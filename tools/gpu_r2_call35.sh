#!/bin/bash
mkdir -p gpurun_out
{
echo "== base (item 4096)"; python tools/slos_profile.py 5 1 4
for v in it2k it8k it16k; do echo "== variant $v"; LWB_LIB_PATH=lightworks_b200/variants/liblwb200_$v.so python tools/slos_profile.py 5 1 4; done
} > gpurun_out/c35_slos.log 2>&1
cat gpurun_out/c35_slos.log
