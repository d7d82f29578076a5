#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:perm_single -s 3 -c 1 -o gpurun_out/c26_k2_24 -f python tools/k2_profile.py 24 > gpurun_out/c26_ncu24.log 2>&1; tail -1 gpurun_out/c26_ncu24.log
ncu --set full --clock-control none --import-source on -k regex:perm_single -s 3 -c 1 -o gpurun_out/c26_k2_26 -f python tools/k2_profile.py 26 > gpurun_out/c26_ncu26.log 2>&1; tail -1 gpurun_out/c26_ncu26.log
