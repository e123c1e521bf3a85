#!/bin/bash
# Round-2 pass Z (8 GPUs): weak-scaling bench with the bucketed (overlapped) and the single all-reduce.
mkdir -p gpurun_out
for b in 0 1; do
SALT_DP_BUCKETS=$b timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2971$b bench.py --gpus 8 --steps 20 --warmup 5 --no-se50 --no-extra --no-cpu-baseline > gpurun_out/r2z_bench_8gpu_buckets$b.json 2> gpurun_out/r2z_bench_8gpu_buckets$b.err
echo "bench 8gpu buckets=$b rc=$?"
done
for f in gpurun_out/r2z_bench_*.json; do echo $f; grep "^{" $f | head -c 250; echo; done
