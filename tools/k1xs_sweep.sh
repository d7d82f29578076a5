#!/bin/bash
# times the full-basis multiplicity-aware walk for the base library and every experiment build under variants/
mkdir -p gpurun_out
python tools/k1x_quick.py base > gpurun_out/k1xs_sweep.log 2>&1
for f in lightworks_b200/variants/liblwb200_*.so; do
  [ -e "$f" ] || continue
  tag=$(basename $f .so); tag=${tag#liblwb200_}
  LWB_LIB_PATH=$PWD/$f python tools/k1x_quick.py $tag >> gpurun_out/k1xs_sweep.log 2>&1
done
cat gpurun_out/k1xs_sweep.log
