#!/bin/bash
# round 2, GPU call 5: per-layer times of both SLOS kernels on C3 + full ncu capture of the two largest K3u layers
mkdir -p gpurun_out
export C3_ONLY=1
python tools/slos_profile.py 1 1 1 > gpurun_out/r2_slos_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:slos -c 40 --csv --log-file gpurun_out/r2_slos_units_launches.csv python tools/slos_profile.py 1 1 1 > /dev/null 2>&1
echo "ncu1 exit=$?"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -k regex:slos -c 40 --csv --log-file gpurun_out/r2_slos_gather_launches.csv python tools/slos_profile.py 1 1 0 > /dev/null 2>&1
echo "ncu2 exit=$?"
ncu --set full --clock-control none --import-source on -k regex:slos_units -s 18 -c 2 -o gpurun_out/r2_prof_slos_units_c3 -f python tools/slos_profile.py 1 1 1 > gpurun_out/r2_slos_ncu.log 2>&1
echo "ncu3 exit=$?"; tail -2 gpurun_out/r2_slos_ncu.log
