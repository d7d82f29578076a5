#!/bin/bash
# round 2, GPU call 9: K3u v3 (async parent block, pipelined stream loads) parity + timing + launch metrics; Batch hook
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "slos or batch or comm or sampler or c3" > gpurun_out/r2_pytest_gpu9.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2_pytest_gpu9.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_gpu9.log | tail -10
python tools/slos_profile.py 3 1 1 2>&1 | tee gpurun_out/r2_slos_timing_v3.log
export C3_ONLY=1
python tools/slos_profile.py 1 1 1 > gpurun_out/r2_slos_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:slos -c 40 --csv --log-file gpurun_out/r2_slos_units_v3_launches.csv python tools/slos_profile.py 1 1 1 > /dev/null 2>&1
echo "ncu exit=$?"
