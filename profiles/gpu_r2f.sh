#!/bin/bash
# Round-2 pass F: CTA-pair convolution (cta_group::2), 256-bit epilogue stores, parallel BatchNorm finalize, segmented backward.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2f_pytest_conv.log 2>&1
echo "pytest conv rc=$?" > gpurun_out/r2f_summary.txt
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_conv_tc_gpu.py > gpurun_out/r2f_pytest_rest.log 2>&1
echo "pytest rest rc=$?" >> gpurun_out/r2f_summary.txt
timeout 300 python bench.py --no-cpu-baseline --no-se50 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?" >> gpurun_out/r2f_summary.txt
SALT_TC_PAIR=0 timeout 300 python bench.py --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2f_bench_nopair.json 2> gpurun_out/r2f_bench_nopair.err
echo "bench nopair rc=$?" >> gpurun_out/r2f_summary.txt
cat gpurun_out/r2f_summary.txt; tail -5 gpurun_out/r2f_pytest_conv.log; tail -8 gpurun_out/r2f_pytest_rest.log
