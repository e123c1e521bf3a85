#!/bin/bash
# Round-2 pass AD (1 GPU): final regression pass on the committed tree: full GPU suite, smoke, bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2ad_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2ad_summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ad_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2ad_summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err
echo "bench rc=$?" >> gpurun_out/r2ad_summary.txt
cat gpurun_out/r2ad_summary.txt; tail -3 gpurun_out/r2ad_pytest.log; tail -2 gpurun_out/r2ad_smoke.log; head -c 900 gpurun_out/r2ad_bench.json
