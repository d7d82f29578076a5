#!/bin/bash
# source-level ncu pass of the shipped multiplicity-aware kernel (128 x 3 launch shape) on C3
mkdir -p gpurun_out
python tools/k1x_profile.py > gpurun_out/c21_plain.log 2>&1; tail -3 gpurun_out/c21_plain.log
ncu --set full --clock-control none --import-source on -k regex:k1x_fused -s 1 -c 1 -o gpurun_out/c21_k1x -f python tools/k1x_profile.py > gpurun_out/c21_ncu.log 2>&1; tail -2 gpurun_out/c21_ncu.log
