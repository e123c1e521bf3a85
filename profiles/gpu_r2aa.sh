#!/bin/bash
# Round-2 pass AA (8 GPUs): what the 120.7 MB gradient all-reduce itself costs under NCCL variants.
mkdir -p gpurun_out
run() { env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT profiles/allreduce_bench.py 2>&1 | grep "^world" ; }
PORT=29801 run >> gpurun_out/r2aa_allreduce.txt
PORT=29802 run NCCL_ALGO=Ring >> gpurun_out/r2aa_allreduce.txt
PORT=29803 run NCCL_ALGO=NVLS >> gpurun_out/r2aa_allreduce.txt
PORT=29804 run NCCL_MIN_NCHANNELS=32 >> gpurun_out/r2aa_allreduce.txt
PORT=29805 run NCCL_ALGO=Tree >> gpurun_out/r2aa_allreduce.txt
cat gpurun_out/r2aa_allreduce.txt
