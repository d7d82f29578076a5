#!/bin/bash
# round 2, GPU call 1: regression tests, accuracy adjudication, cmul-form A/B, ncu record of the DFMA peak probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2_pytest_gpu.log
tail -4 gpurun_out/r2_pytest_gpu.log
timeout 600 python tools/accuracy_probe.py > gpurun_out/r2_accuracy_probe.log 2>&1; echo "probe exit=$?"
cat gpurun_out/r2_accuracy_probe.log
export NS=12,16,20
EXTRA=none timeout 300 python tools/k1_timing.py base > gpurun_out/r2_k1_base.log 2>&1
for t in cm1 ch3cm1; do
  LWB_LIB_PATH=$PWD/lightworks_b200/variants/liblwb200_$t.so EXTRA=none timeout 300 python tools/k1_timing.py $t > gpurun_out/r2_k1_$t.log 2>&1
done
grep -h "frac" gpurun_out/r2_k1_*.log
python tools/fp64_probe.py > gpurun_out/r2_fp64_probe.log 2>&1 && \
ncu --set full --clock-control none -k regex:dfma_peak -c 2 --csv --page raw --log-file gpurun_out/r2_dfma_peak_ncu.csv python tools/fp64_probe.py > gpurun_out/r2_dfma_ncu.log 2>&1
echo "ncu exit=$?"; tail -3 gpurun_out/r2_fp64_probe.log
