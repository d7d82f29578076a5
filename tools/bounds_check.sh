#!/bin/bash
# compute-sanitizer is closed on the measurement pool: run every kernel family on the small configurations against a
# build with device-side index asserts (make TAG=bounds EXTRA=-DLWB_DEBUG_BOUNDS -> variants/liblwb200_bounds.so).
cd "$(dirname "$0")/.."
LWB_LIB_PATH=$PWD/lightworks_b200/variants/liblwb200_bounds.so python tools/sanitize_small.py
