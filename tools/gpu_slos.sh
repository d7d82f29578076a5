#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -x -k "slos or pdist or basis_probs or c3" 2>&1 | tail -3
CMD="python tools/slos_profile.py 2"
$CMD > gpurun_out/slos_plain.log 2>&1 && cat gpurun_out/slos_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:slos_layer -s 19 -c 1 -o gpurun_out/prof_slos -f $CMD > gpurun_out/slos_ncu_full.log 2>&1
echo "ncu full exit=$?"
