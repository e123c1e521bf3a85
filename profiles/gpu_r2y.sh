#!/bin/bash
# Round-2 pass Y (2 GPUs): the NCCL data-parallel parity test, the bench with bucketed (overlapped) and single all-reduce.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2y_pytest_dp.log 2>&1
echo "pytest dp rc=$?" > gpurun_out/r2y_summary.txt
for b in 1 0; do
SALT_DP_BUCKETS=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$b bench.py --gpus 2 --steps 20 --warmup 5 --no-se50 --no-extra > gpurun_out/r2y_bench_2gpu_buckets$b.json 2> gpurun_out/r2y_bench_2gpu_buckets$b.err
echo "bench 2gpu buckets=$b rc=$?" >> gpurun_out/r2y_summary.txt
done
cat gpurun_out/r2y_summary.txt; grep -E "step 0|after 4|rank . losses|passed|failed" gpurun_out/r2y_pytest_dp.log; for f in gpurun_out/r2y_bench_*.json; do echo $f; head -c 250 $f; echo; done
