#!/bin/bash
# Round-2 pass G (2 GPUs): 2-rank NCCL data-parallel parity test, 2-GPU bench (bucketed all-reduce overlapped with backward); plus the
# conv / engine suites for the 32-channel row-halo path and the shuffle block reduction.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2g_pytest_dp.log 2>&1
echo "pytest dp rc=$?" > gpurun_out/r2g_summary.txt
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
echo "pytest conv+engine rc=$?" >> gpurun_out/r2g_summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-se50 > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err
echo "bench 2gpu rc=$?" >> gpurun_out/r2g_summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 > gpurun_out/r2g_bench_1gpu.json 2> gpurun_out/r2g_bench_1gpu.err
echo "bench 1gpu rc=$?" >> gpurun_out/r2g_summary.txt
cat gpurun_out/r2g_summary.txt; tail -12 gpurun_out/r2g_pytest_dp.log; tail -3 gpurun_out/r2g_pytest.log; head -c 400 gpurun_out/r2g_bench_2gpu.json
