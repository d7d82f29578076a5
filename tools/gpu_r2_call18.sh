#!/bin/bash
# round 2, GPU call 18: negated-table trick (dynamic-sign DFMA -> DADD): tests + timings of K1, K1x, K2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu18.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2_pytest_gpu18.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_gpu18.log | tail -10
python tools/k1x_quick.py negtab 2>&1 | tee gpurun_out/r2_k1x_negtab.log
NS=12,16,20 EXTRA=C3 timeout 600 python tools/k1_timing.py negtab 2>&1 | tee gpurun_out/r2_k1_negtab.log
