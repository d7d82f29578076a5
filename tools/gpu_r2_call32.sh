#!/bin/bash
# gpurun --gpus 2: communicator tests + N=2 bench with the C5 extras after the small-permanent rule
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_comm.py -m gpu -q 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c32_bench_n2.json 2> gpurun_out/c32_bench_n2.err; echo "bench exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/c32_bench_n2.json')); print(d['ms_per_step'])
for k,v in d['extra']['single_perm_sharded'].items(): print(k, round(v['ms']*1e3,1),'us vs one GPU', round(v['ms_one_gpu']*1e3,1), v['path'][:30])
PY
