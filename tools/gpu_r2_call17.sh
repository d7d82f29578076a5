#!/bin/bash
# K3u record kernel: variant sweep (children per thread, prefetch depth, CTAs/SM, unit size), then ncu of the base build
mkdir -p gpurun_out
{
python tools/slos_profile.py 5 1 1
echo "== base (J2 PF2 TMAX 8192)"; python tools/slos_profile.py 5 1 4
for v in a b c d; do echo "== variant $v"; LWB_LIB_PATH=lightworks_b200/variants/liblwb200_$v.so python tools/slos_profile.py 5 1 4; done
} > gpurun_out/c17_slos.log 2>&1
cat gpurun_out/c17_slos.log
C3_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:slos_units4 -s 29 -c 1 -o gpurun_out/c17_su4 -f python tools/slos_profile.py 2 1 4 > gpurun_out/c17_ncu.log 2>&1
tail -2 gpurun_out/c17_ncu.log
