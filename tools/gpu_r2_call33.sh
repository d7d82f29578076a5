#!/bin/bash
# gpurun --gpus 8: N = 8 bench line of the final build (C5 extras with the small-permanent rule)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench exit=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r2_bench_n8.json')); print(d['ms_per_step'], d['per_rank_compute_ms'])
for k,v in d['extra']['single_perm_sharded'].items(): print(k, round(v['ms']*1e3,1),'us vs one GPU', round(v['ms_one_gpu']*1e3,1), round(v['speedup'],2))
PY
