#!/bin/bash
# round 2, GPU call 12: bench N=1 (both arms), compute-sanitizer memcheck on the small configurations
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_n1_full.json 2> gpurun_out/r2_bench_n1_full.err; echo "bench exit=$?"; cut -c1-1500 gpurun_out/r2_bench_n1_full.json; tail -3 gpurun_out/r2_bench_n1_full.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref exit=$?"; cut -c1-600 gpurun_out/r2_bench_ref.json
python tools/sanitize_small.py > gpurun_out/r2_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --leak-check full --log-file gpurun_out/r2_sanitizer_memcheck.log python tools/sanitize_small.py > gpurun_out/r2_sanitize_memcheck_stdout.log 2>&1
echo "memcheck exit=$?"; tail -5 gpurun_out/r2_sanitizer_memcheck.log; tail -2 gpurun_out/r2_sanitize_memcheck_stdout.log
