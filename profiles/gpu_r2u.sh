#!/bin/bash
# Round-2 pass U (1 GPU): upsample backward with several source rows per block, scSE apply on 16 consumer warps, fast sigmoid - parity suites, timeline, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py tests/test_dropin_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2u_summary.txt
timeout 200 python profiles/step_timeline.py > gpurun_out/r2u_step_timeline.txt 2>&1
echo "timeline rc=$?" >> gpurun_out/r2u_summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
echo "bench rc=$?" >> gpurun_out/r2u_summary.txt
cat gpurun_out/r2u_summary.txt; tail -3 gpurun_out/r2u_pytest.log; head -1 gpurun_out/r2u_step_timeline.txt; head -c 300 gpurun_out/r2u_bench.json
