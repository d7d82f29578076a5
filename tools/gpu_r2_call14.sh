#!/bin/bash
# round 2, GPU call 14: source-level ncu capture of the K1x fused kernel on C3
mkdir -p gpurun_out
python tools/k1x_profile.py > gpurun_out/r2_k1x_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1x_fused -s 1 -c 1 -o gpurun_out/r2_prof_k1x -f python tools/k1x_profile.py > gpurun_out/r2_k1x_ncu.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/r2_k1x_plain.log
