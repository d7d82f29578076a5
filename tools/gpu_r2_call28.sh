#!/bin/bash
# round-2 ncu launch list of the bench command (after the same command exited 0 without ncu)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/c28_bench.json 2> gpurun_out/c28_bench.err; echo "bench exit=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c28_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/c28_ncu.log 2>&1; echo "ncu exit=$?"; wc -l gpurun_out/c28_launches.csv
