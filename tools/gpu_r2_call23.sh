#!/bin/bash
# SLOS kernels (gather, unit, record) with device-side index asserts (-DLWB_DEBUG_BOUNDS build of slos_units.cu)
mkdir -p gpurun_out
export LWB_LIB_PATH=$PWD/lightworks_b200/variants/liblwb200_bounds.so
{
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "slos" 2>&1 | tail -3
python tools/sanitize_small.py 2>&1 | tail -3
C3_ONLY=1 python tools/slos_profile.py 1 1 4
C3_ONLY=1 python tools/slos_profile.py 1 0 4
} > gpurun_out/c23_bounds.log 2>&1
cat gpurun_out/c23_bounds.log
