#!/bin/bash
# round 2, GPU call 4: tests (K3u, fused K2 reduction, comm fix), SLOS and K2 timings, ncu of the new SLOS kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu4.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2_pytest_gpu4.log
tail -25 gpurun_out/r2_pytest_gpu4.log
python tools/slos_profile.py 3 1 0 2>&1 | tee gpurun_out/r2_slos_timing.log
python tools/slos_profile.py 3 1 1 2>&1 | tee -a gpurun_out/r2_slos_timing.log
NS=12 EXTRA=C3 timeout 300 python tools/k1_timing.py k2fused 2>&1 | grep single | tee gpurun_out/r2_k2_fused.log
python tools/slos_profile.py 1 1 1 > gpurun_out/r2_slos_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:slos_units -s 27 -c 3 -o gpurun_out/r2_prof_slos_units -f python tools/slos_profile.py 1 1 1 > gpurun_out/r2_slos_ncu.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/r2_slos_ncu.log
