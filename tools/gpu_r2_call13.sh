#!/bin/bash
# round 2, GPU call 13 (gpurun --gpus 8): bench through the communicator at N = 8 and N = 4, 2-process + LOCAL comm tests
mkdir -p gpurun_out
for N in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N"
  EX=""; [ $N != 8 ] && EX="--no-extras"
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 $EX > gpurun_out/r2_bench_n${N}.json 2> gpurun_out/r2_bench_n${N}.err; echo "bench N=$N exit=$?"; cut -c1-700 gpurun_out/r2_bench_n${N}.json; tail -3 gpurun_out/r2_bench_n${N}.err
done
LWB_BENCH_PEER_STORES=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_n8_nccl.json 2> gpurun_out/r2_bench_n8_nccl.err; echo "bench N=8 nccl exit=$?"; cut -c1-400 gpurun_out/r2_bench_n8_nccl.json
timeout 900 python -m pytest tests/test_gpu_comm.py -m gpu -q > gpurun_out/r2_pytest_comm8.log 2>&1; tail -4 gpurun_out/r2_pytest_comm8.log
