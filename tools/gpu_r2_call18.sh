#!/bin/bash
# shipped K3r record kernel: GPU suite, then per-layer ncu timing list and one full capture of the last layer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c18_pytest.log 2>&1; echo "pytest exit=$?"; tail -3 gpurun_out/c18_pytest.log
for u in 0 1 4; do C3_ONLY=1 python tools/slos_profile.py 5 1 $u; done > gpurun_out/c18_slos.log 2>&1; cat gpurun_out/c18_slos.log
C3_ONLY=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:slos -c 40 --csv --log-file gpurun_out/c18_layers.csv python tools/slos_profile.py 1 1 4 > gpurun_out/c18_ncu1.log 2>&1
C3_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:slos_units4 -s 19 -c 1 -o gpurun_out/c18_su4 -f python tools/slos_profile.py 1 1 4 > gpurun_out/c18_ncu2.log 2>&1
tail -2 gpurun_out/c18_ncu2.log
