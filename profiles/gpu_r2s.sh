#!/bin/bash
# Round-2 pass S (1 GPU): halo-box weight-gradient kernel (3x3 stride 1) + the upsample-backward table fix: full GPU suite,
# timeline, bench with the halo kernel on / off.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2s_summary.txt
timeout 200 python profiles/step_timeline.py > gpurun_out/r2s_step_timeline.txt 2>&1
echo "timeline rc=$?" >> gpurun_out/r2s_summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?" >> gpurun_out/r2s_summary.txt
SALT_WGRAD_HALO=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2s_bench_nohalo.json 2> gpurun_out/r2s_bench_nohalo.err
echo "bench nohalo rc=$?" >> gpurun_out/r2s_summary.txt
cat gpurun_out/r2s_summary.txt; tail -5 gpurun_out/r2s_pytest.log; head -1 gpurun_out/r2s_step_timeline.txt; head -c 300 gpurun_out/r2s_bench.json; echo; head -c 300 gpurun_out/r2s_bench_nohalo.json
