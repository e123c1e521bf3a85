#!/bin/bash
# Round-2 pass X (1 GPU): resident-weight row-halo convolution (<= 64 input channels, one channel tile): conv + engine parity,
# timeline, bench on / off; racecheck with and without the CTA-pair kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_conv_tc_gpu.py tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2x_summary.txt
timeout 200 python profiles/step_timeline.py > gpurun_out/r2x_step_timeline.txt 2>&1
echo "timeline rc=$?" >> gpurun_out/r2x_summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
echo "bench rc=$?" >> gpurun_out/r2x_summary.txt
SALT_TC_RESW=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2x_bench_noresw.json 2> gpurun_out/r2x_bench_noresw.err
echo "bench noresw rc=$?" >> gpurun_out/r2x_summary.txt
SALT_TC_PAIR=0 timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python profiles/sanitize_step.py > gpurun_out/r2x_racecheck_nopair.log 2>&1
echo "racecheck nopair rc=$?" >> gpurun_out/r2x_summary.txt
cat gpurun_out/r2x_summary.txt; tail -3 gpurun_out/r2x_pytest.log; head -1 gpurun_out/r2x_step_timeline.txt; head -c 300 gpurun_out/r2x_bench.json; echo; head -c 300 gpurun_out/r2x_bench_noresw.json; echo; tail -2 gpurun_out/r2x_racecheck_nopair.log
