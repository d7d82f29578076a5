#!/bin/bash
# Round-end evidence: ncu launch list and one full capture of the dominant kernel for bench.py's own command.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras --cpu-sample 2000"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k1x_fused -s 3 -c 1 -o gpurun_out/prof_bench_k1x -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"
