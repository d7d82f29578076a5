#!/bin/bash
# runs k1_timing.py for the base library and every experiment build under lightworks_b200/variants
mkdir -p gpurun_out
export NS=${NS:-8,10,12,14,16,20}
python tools/k1_timing.py base > gpurun_out/timing_base.log 2>&1
for f in lightworks_b200/variants/liblwb200_*.so; do
  [ -e "$f" ] || continue
  tag=$(basename $f .so); tag=${tag#liblwb200_}
  LWB_LIB_PATH=$PWD/$f EXTRA=C3 python tools/k1_timing.py $tag > gpurun_out/timing_$tag.log 2>&1
done
