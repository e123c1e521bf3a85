#!/bin/bash
# Round-2 pass C: pipeline-stall accounting of the row-halo convolution (timing build), GPU suite after the split-accumulator fix.
mkdir -p gpurun_out
SALT_LIB_PATH=open-solution-salt-identification_b200/libsaltunet_timing.so timeout 120 python profiles/rows_timing.py > gpurun_out/r2c_rows_timing.txt 2>&1
echo "rows_timing rc=$?" > gpurun_out/r2c_summary.txt
SALT_TC_CLUSTER=1 SALT_LIB_PATH=open-solution-salt-identification_b200/libsaltunet_timing.so timeout 120 python profiles/rows_timing.py > gpurun_out/r2c_rows_timing_cl1.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_summary.txt
timeout 600 python -m pytest tests/test_baseline_configs_gpu.py tests/test_engine_gpu.py tests/test_conv_tc_gpu.py -k "config or fused or split or forward_eval" -m gpu -q -p no:cacheprovider -s 2>&1 | grep -E "config|eval forward|split-bf16 tc conv fwd|eval logits|trained|passed|failed" > gpurun_out/r2c_pytest_verbose.log
cat gpurun_out/r2c_summary.txt; cat gpurun_out/r2c_rows_timing.txt; tail -25 gpurun_out/r2c_pytest_gpu.log
