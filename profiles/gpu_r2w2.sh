#!/bin/bash
# Round-2 pass W2 (1 GPU): record pass on the final convolution sources (after the resident-weight kernel): full suite, smoke, bench
# line, per-launch timeline, ncu launch list with DRAM bytes.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2w2_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2w2_summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w2_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2w2_summary.txt
timeout 900 python bench.py > gpurun_out/r2w2_bench.json 2> gpurun_out/r2w2_bench.err
echo "bench rc=$?" >> gpurun_out/r2w2_summary.txt
timeout 200 python profiles/step_timeline.py > gpurun_out/r2w2_step_timeline.txt 2>&1
echo "timeline rc=$?" >> gpurun_out/r2w2_summary.txt
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2w2_launches.csv python profiles/one_step.py 2 > gpurun_out/r2w2_one_step_under_ncu.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/r2w2_summary.txt
cat gpurun_out/r2w2_summary.txt; tail -3 gpurun_out/r2w2_pytest.log; tail -2 gpurun_out/r2w2_smoke.log; head -c 400 gpurun_out/r2w2_bench.json
