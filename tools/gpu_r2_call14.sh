#!/bin/bash
# K3u v4 (sub-units): bit-identity against the gather kernel, then timing v3 vs v4
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "slos" > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c14_pytest.log
tail -5 gpurun_out/c14_pytest.log
for u in 1 4 1 4; do python tools/slos_profile.py 5 1 $u; done > gpurun_out/c14_slos.log 2>&1
cat gpurun_out/c14_slos.log
