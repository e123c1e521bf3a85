#!/bin/bash
# Round-2 pass I (1 GPU): 4 accumulator buffers for the narrow tiles, prefetched TMEM loads for the wide ones, multiply-shift tile
# index maths, bias of the narrow epilogue in shared memory: parity suites, bench, pipeline-stall accounting.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2i_summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$?" >> gpurun_out/r2i_summary.txt
SALT_LIB_PATH=open-solution-salt-identification_b200/libsaltunet_timing.so timeout 300 python profiles/rows_timing.py > gpurun_out/r2i_rows_timing.txt 2>&1
echo "timing rc=$?" >> gpurun_out/r2i_summary.txt
cat gpurun_out/r2i_summary.txt; tail -3 gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_rows_timing.txt; head -c 300 gpurun_out/r2i_bench.json
