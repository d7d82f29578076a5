#!/bin/bash
# K2 with odd row stride: parity tests of single permanents, timings n = 20 .. 30
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "single or c5 or perm_single" 2>&1 | tail -3
python tools/k2_profile.py 20 21 22 23 24 25 26 27 28 30 32 > gpurun_out/c25_k2.log 2>&1; cat gpurun_out/c25_k2.log
