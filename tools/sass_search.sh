#!/bin/bash
# Offline screen of source-level variants of the Glynn inner loop: compiles ONE instantiation per variant and
# prints the modelled FP64 issue cycles of its hot loop (tools/sass_fp64_model.py).  No GPU needed.
# usage: tools/sass_search.sh N LOG_L "<flags variant 1>" "<flags variant 2>" ...
cd "$(dirname "$0")/.."
N=$1; L=$2; shift 2
for V in "$@"; do
  O=/tmp/lab/v_$(echo "$V" | tr -c 'A-Za-z0-9' '_').o
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Ilightworks_b200/csrc -DLAB_N=$N -DLAB_L=$L $V -c /tmp/lab/one.cu -o $O 2>/dev/null || { echo "FAILED $V"; continue; }
  echo "== $V  regs: $(cuobjdump -res-usage $O 2>/dev/null | grep -o 'REG:[0-9]*' | head -1)"
  cuobjdump -sass $O | python tools/sass_fp64_model.py perm_matrices | tail -2
done
