#!/bin/bash
# ncu evidence for the C4 source-statistics combination: launch list of the convolution kernel and one full capture
mkdir -p gpurun_out
CMD="python tools/c4_end_to_end.py"
$CMD > gpurun_out/c4_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:convolve_kernel --csv --log-file gpurun_out/c4_convolve_launches.csv $CMD > gpurun_out/c4_ncu_launches.log 2>&1
echo "ncu launches exit=$?"
ncu --set full --clock-control none --import-source on -k regex:convolve_kernel -s 700 -c 1 -o gpurun_out/prof_c4_convolve -f $CMD > gpurun_out/c4_ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -2 gpurun_out/c4_ncu_full.log
