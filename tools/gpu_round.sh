#!/bin/bash
# One GPU session: tests, bench (both arms), then the ncu launch list and one full capture of K1.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
CMD="python bench.py --steps 2 --warmup 3 --no-extras --cpu-sample 2000"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:perm_states -s 3 -c 1 -o gpurun_out/prof_k1 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -3 gpurun_out/ncu_full.log
