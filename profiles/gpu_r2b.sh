#!/bin/bash
# Round-2 pass B: fixed operand-feed micro-benchmark; full GPU suite on the new code (split-bf16 fp32 mode on tcgen05, fused eval
# epilogues, wide N tiles, deterministic statistics); bench with and without the wide tiles.
mkdir -p gpurun_out
timeout 90 profiles/_bin/mma_feed > gpurun_out/r2b_mma_feed.txt 2>&1
echo "mma_feed rc=$?" > gpurun_out/r2b_summary.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2b_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_summary.txt
timeout 600 python -m pytest tests/test_baseline_configs_gpu.py tests/test_engine_gpu.py -k "config or fused or reproducible" -m gpu -q -p no:cacheprovider -s > gpurun_out/r2b_pytest_verbose.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-se50 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?" >> gpurun_out/r2b_summary.txt
SALT_TC_WIDE=0 timeout 300 python bench.py --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2b_bench_nowide.json 2> gpurun_out/r2b_bench_nowide.err
echo "bench nowide rc=$?" >> gpurun_out/r2b_summary.txt
cat gpurun_out/r2b_summary.txt; tail -15 gpurun_out/r2b_pytest_gpu.log; cat gpurun_out/r2b_mma_feed.txt
