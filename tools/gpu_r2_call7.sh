#!/bin/bash
# round 2, GPU call 7: full GPU suite after the lazy-result / vectorised-runner changes, drop-in timings, K3u v2 timings + ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_gpu7.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2_pytest_gpu7.log
grep -E "passed|failed|FAILED|Error|C3 Sampler" gpurun_out/r2_pytest_gpu7.log | tail -20
timeout 900 python tools/dropin_bench.py > gpurun_out/r2_dropin_bench.log 2>&1; echo "dropin exit=$?"; grep "^{" gpurun_out/r2_dropin_bench.log | cut -c1-260
python tools/slos_profile.py 3 1 1 2>&1 | tee gpurun_out/r2_slos_timing_v2.log
export C3_ONLY=1
python tools/slos_profile.py 1 1 1 > gpurun_out/r2_slos_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:slos -c 40 --csv --log-file gpurun_out/r2_slos_units_v2_launches.csv python tools/slos_profile.py 1 1 1 > /dev/null 2>&1
echo "ncu exit=$?"
