#!/bin/bash
# K3u v4: parity, timing, then one ncu --set full capture of the last layer (largest launch)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slos" > gpurun_out/c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c15_pytest.log
tail -3 gpurun_out/c15_pytest.log
for u in 1 4; do C3_ONLY=1 python tools/slos_profile.py 5 1 $u; done > gpurun_out/c15_slos.log 2>&1
cat gpurun_out/c15_slos.log
C3_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:slos_units4 -s 29 -c 1 -o gpurun_out/c15_su4 -f python tools/slos_profile.py 2 1 4 > gpurun_out/c15_ncu.log 2>&1
tail -3 gpurun_out/c15_ncu.log
