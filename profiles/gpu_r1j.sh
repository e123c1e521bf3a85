#!/bin/bash
# parity of the experimental multi-sub-tile convolution (uses whatever GPU budget is left)
mkdir -p gpurun_out
SALT_TC_MULTI=1 SALT_TC_CLUSTER=1 timeout 70 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x -p no:cacheprovider -k forward_and_dgrad > gpurun_out/pytest_multi_j.log 2>&1
echo "pytest conv (multi) rc=$?" > gpurun_out/summary_j.txt
tail -12 gpurun_out/pytest_multi_j.log; cat gpurun_out/summary_j.txt
SALT_TC_MULTI=1 SALT_TC_CLUSTER=1 timeout 60 python bench.py --no-extra --no-cpu-baseline --steps 5 > gpurun_out/bench_multi.json 2> gpurun_out/bench_multi.err
head -c 600 gpurun_out/bench_multi.json
