#!/bin/bash
# 2-GPU check: smoke on GPU 0, then the data-parallel bench (CUDA-graph replay per rank + eager NCCL all-reduce).
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_f.log 2>&1
echo "smoke rc=$?" > gpurun_out/summary_f.txt
tail -4 gpurun_out/smoke_f.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra \
  > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench 2gpu rc=$?" >> gpurun_out/summary_f.txt
cat gpurun_out/summary_f.txt; head -c 900 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
