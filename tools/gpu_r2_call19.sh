#!/bin/bash
# gpurun --gpus 8: scaling series after the refit of the per-state overhead of the equal-work cuts
mkdir -p gpurun_out
for N in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N"
  EX=""; [ $N != 8 ] && EX="--no-extras"
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 $EX > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench N=$N exit=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n${N}.json')); print(d['n_gpus'], d['ms_per_step'], d['per_rank_compute_ms'])"
done
