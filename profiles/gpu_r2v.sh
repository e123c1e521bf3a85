#!/bin/bash
# Round-2 pass V (1 GPU): N = 128 row-variant of the halo weight gradient: conv + engine parity, timeline, bench on / off.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_conv_tc_gpu.py tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2v_summary.txt
timeout 200 python profiles/step_timeline.py > gpurun_out/r2v_step_timeline.txt 2>&1
echo "timeline rc=$?" >> gpurun_out/r2v_summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
echo "bench rc=$?" >> gpurun_out/r2v_summary.txt
SALT_WGRAD_HALO_ROWS=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2v_bench_norows.json 2> gpurun_out/r2v_bench_norows.err
echo "bench norows rc=$?" >> gpurun_out/r2v_summary.txt
cat gpurun_out/r2v_summary.txt; tail -5 gpurun_out/r2v_pytest.log; head -1 gpurun_out/r2v_step_timeline.txt; head -c 300 gpurun_out/r2v_bench.json; echo; head -c 300 gpurun_out/r2v_bench_norows.json
