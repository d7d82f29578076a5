#!/bin/bash
# round 2, GPU call 6: full GPU suite (array sampler, drop-in incl. the reference's own emulator tests), drop-in timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/r2_pytest_gpu6.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2_pytest_gpu6.log
grep -E "passed|failed|FAILED|Error|C3 Sampler" gpurun_out/r2_pytest_gpu6.log | tail -20
timeout 900 python tools/dropin_bench.py > gpurun_out/r2_dropin_bench.log 2>&1; echo "dropin exit=$?"; grep -v "^\[" gpurun_out/r2_dropin_bench.log | tail -20
