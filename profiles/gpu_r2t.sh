#!/bin/bash
# Round-2 pass T (1 GPU): multiply-shift divisions in the ring ops, halo wgrad only for <= 64 output channels with rotated epilogue order - parity suites, timeline, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py tests/test_dropin_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2t_summary.txt
timeout 200 python profiles/step_timeline.py > gpurun_out/r2t_step_timeline.txt 2>&1
echo "timeline rc=$?" >> gpurun_out/r2t_summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
echo "bench rc=$?" >> gpurun_out/r2t_summary.txt
cat gpurun_out/r2t_summary.txt; tail -3 gpurun_out/r2t_pytest.log; head -1 gpurun_out/r2t_step_timeline.txt; head -c 300 gpurun_out/r2t_bench.json
