#!/bin/bash
# Round-2 pass H (2 GPUs): 2-rank NCCL parity test (restructured), bucketed vs single all-reduce at 2 GPUs, ncu --set full of the
# row-halo / CTA-pair convolutions (source-level stall reasons of the epilogue).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider -s -k two_ranks > gpurun_out/r2h_pytest_dp.log 2>&1
echo "pytest dp rc=$?" > gpurun_out/r2h_summary.txt
for b in 1 0; do
SALT_DP_BUCKETS=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$b bench.py --gpus 2 --steps 20 --warmup 5 --no-se50 --no-extra > gpurun_out/r2h_bench_2gpu_buckets$b.json 2> gpurun_out/r2h_bench_2gpu_buckets$b.err
echo "bench 2gpu buckets=$b rc=$?" >> gpurun_out/r2h_summary.txt
done
NCCL_MAX_NCHANNELS=4 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-se50 --no-extra > gpurun_out/r2h_bench_2gpu_nch4.json 2> gpurun_out/r2h_bench_2gpu_nch4.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2h_bench_1gpu.json 2> gpurun_out/r2h_bench_1gpu.err
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_rows -c 6 -f -o gpurun_out/r2h_rows_full python profiles/microbench_conv.py > gpurun_out/r2h_ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r2h_summary.txt
cat gpurun_out/r2h_summary.txt; grep -E "step 0|after 4|rank . losses|passed|failed" gpurun_out/r2h_pytest_dp.log; for f in gpurun_out/r2h_bench_*.json; do echo $f; head -c 250 $f; echo; done
