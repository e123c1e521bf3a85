#!/bin/bash
# Round-2 pass Q (1 GPU): 16 consumer warps in the ring ops that fit, tabulated weights in the fused upsample backward, shorter index
# maths in fold backward: full GPU suite, timeline, bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2q_summary.txt
timeout 200 python profiles/step_timeline.py > gpurun_out/r2q_step_timeline.txt 2>&1
echo "timeline rc=$?" >> gpurun_out/r2q_summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-se50 --no-extra > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
echo "bench rc=$?" >> gpurun_out/r2q_summary.txt
cat gpurun_out/r2q_summary.txt; tail -3 gpurun_out/r2q_pytest.log; head -1 gpurun_out/r2q_step_timeline.txt; head -c 300 gpurun_out/r2q_bench.json
