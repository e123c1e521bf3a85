#!/bin/bash
# Round-2 pass A: tcgen05 operand-feed micro-benchmark, full GPU test suite (incl. the new BASELINE-config + reproducibility tests),
# smoke, bench, racecheck of one training step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt
timeout 300 profiles/_bin/mma_feed > gpurun_out/r2a_mma_feed.txt 2>&1
echo "mma_feed rc=$?" > gpurun_out/r2a_summary.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_summary.txt
timeout 1500 python -m pytest tests/test_baseline_configs_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2a_pytest_baseline.log 2>&1
echo "pytest baseline rc=$?" >> gpurun_out/r2a_summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2a_summary.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_summary.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python profiles/sanitize_step.py > gpurun_out/r2a_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2a_summary.txt
cat gpurun_out/r2a_summary.txt; tail -5 gpurun_out/r2a_pytest_gpu.log; tail -40 gpurun_out/r2a_mma_feed.txt; tail -5 gpurun_out/r2a_racecheck.log
