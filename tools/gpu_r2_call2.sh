#!/bin/bash
# round 2, GPU call 2 (gpurun --gpus 2): full GPU test-suite incl. the communicators, multi-GPU bench through the C ABI
mkdir -p gpurun_out
N=${N:-2}
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2_pytest_gpu2.log
tail -15 gpurun_out/r2_pytest_gpu2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench N=$N exit=$?"; cat gpurun_out/r2_bench_n${N}.json; tail -5 gpurun_out/r2_bench_n${N}.err
LWB_BENCH_PEER_STORES=0 timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_n${N}_nccl.json 2> gpurun_out/r2_bench_n${N}_nccl.err; echo "bench N=$N nccl exit=$?"; cat gpurun_out/r2_bench_n${N}_nccl.json; tail -5 gpurun_out/r2_bench_n${N}_nccl.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --cpu-sample 20000 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench N=1 exit=$?"; cat gpurun_out/r2_bench_n1.json; tail -3 gpurun_out/r2_bench_n1.err
