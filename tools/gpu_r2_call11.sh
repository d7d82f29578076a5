#!/bin/bash
mkdir -p gpurun_out
export C3_ONLY=1
python tools/slos_profile.py 3 1 1 2>&1 | sed 's/^/base     /' | tee gpurun_out/r2_slos_variant_sweep.log
for f in lightworks_b200/variants/liblwb200_*.so; do
  tag=$(basename $f .so); tag=${tag#liblwb200_}
  LWB_LIB_PATH=$PWD/$f python tools/slos_profile.py 3 1 1 2>&1 | sed "s/^/$tag  /" | tee -a gpurun_out/r2_slos_variant_sweep.log
done
