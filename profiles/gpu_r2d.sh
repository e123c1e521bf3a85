#!/bin/bash
# Round-2 pass D: early barrier probes in the three tcgen05 issue loops - stall accounting, parity, bench.
mkdir -p gpurun_out
SALT_LIB_PATH=open-solution-salt-identification_b200/libsaltunet_timing.so timeout 120 python profiles/rows_timing.py > gpurun_out/r2d_rows_timing.txt 2>&1
echo "rows_timing rc=$?" > gpurun_out/r2d_summary.txt
timeout 900 python -m pytest tests/test_conv_tc_gpu.py tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_summary.txt
timeout 300 python bench.py --no-cpu-baseline --no-se50 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?" >> gpurun_out/r2d_summary.txt
cat gpurun_out/r2d_summary.txt; cat gpurun_out/r2d_rows_timing.txt; tail -5 gpurun_out/r2d_pytest.log; head -c 1500 gpurun_out/r2d_bench.json
