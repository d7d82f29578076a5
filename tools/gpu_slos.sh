#!/bin/bash
mkdir -p gpurun_out
python tools/k1_timing.py base > gpurun_out/timing_k2v2.log 2>&1; grep -E "single|C3" gpurun_out/timing_k2v2.log
python -m pytest tests/test_gpu_parity.py -q -k "single or smoke" 2>&1 | tail -3
CMD="python tools/slos_profile.py 2"
$CMD > gpurun_out/slos_plain.log 2>&1 && cat gpurun_out/slos_plain.log && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed.sum,l1tex__t_sector_hit_rate.pct --clock-control none -c 40 --csv --log-file gpurun_out/slos_launches.csv $CMD > gpurun_out/slos_ncu.log 2>&1
echo "ncu exit=$?"
