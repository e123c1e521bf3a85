#!/bin/bash
# Round-2 pass L (1 GPU): BatchNorm passes on the cp.async.bulk stream ring - parity suites, per-launch timeline, bench A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py tests/test_dp_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2l_summary.txt
timeout 200 python profiles/step_timeline.py > gpurun_out/r2l_step_timeline.txt 2>&1
echo "timeline rc=$?" >> gpurun_out/r2l_summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
echo "bench rc=$?" >> gpurun_out/r2l_summary.txt
SALT_EW_RING=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2l_bench_noring.json 2> gpurun_out/r2l_bench_noring.err
echo "bench noring rc=$?" >> gpurun_out/r2l_summary.txt
cat gpurun_out/r2l_summary.txt; tail -3 gpurun_out/r2l_pytest.log; head -1 gpurun_out/r2l_step_timeline.txt; head -c 300 gpurun_out/r2l_bench.json; echo; head -c 300 gpurun_out/r2l_bench_noring.json
