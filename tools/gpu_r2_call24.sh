#!/bin/bash
# K2 at n = 24: event-timed vs back-to-back vs ncu kernel duration (where do the ~35 us over the n = 28 rate go?)
mkdir -p gpurun_out
python tools/k2_profile.py 22 24 26 28 > gpurun_out/c24_k2.log 2>&1; cat gpurun_out/c24_k2.log
ncu --set full --clock-control none --import-source on -k regex:perm_single -s 3 -c 1 -o gpurun_out/c24_k2 -f python tools/k2_profile.py 24 > gpurun_out/c24_ncu.log 2>&1; tail -2 gpurun_out/c24_ncu.log
