#!/bin/bash
# round 2, GPU call 8: GPU suite (Batch across unitaries, __ne__ fix), full ncu of K3u v2 on the last two C3 layers
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_gpu8.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2_pytest_gpu8.log
grep -E "passed|failed|FAILED|Error|C3 Sampler" gpurun_out/r2_pytest_gpu8.log | tail -20
export C3_ONLY=1
python tools/slos_profile.py 1 1 1 > gpurun_out/r2_slos_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:slos_units -s 18 -c 2 -o gpurun_out/r2_prof_slos_units_v2 -f python tools/slos_profile.py 1 1 1 > gpurun_out/r2_slos_ncu.log 2>&1
echo "ncu exit=$?"
