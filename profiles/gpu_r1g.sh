#!/bin/bash
# Last GPU pass of round 1: parity suite of the final tree (incl. UNetSeResNet-101), smoke, then the experimental CTA-pair
# (cta_group::2) row-halo convolution: parity first, bench only if parity is green.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_g.log 2>&1
echo "pytest rc=$?" > gpurun_out/summary_g.txt
tail -3 gpurun_out/pytest_gpu_g.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_g.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary_g.txt
SALT_TC_PAIR=1 timeout 240 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x -p no:cacheprovider -k forward_and_dgrad > gpurun_out/pytest_pair_g.log 2>&1
rc=$?
echo "pytest conv (pair) rc=$rc" >> gpurun_out/summary_g.txt
tail -15 gpurun_out/pytest_pair_g.log
if [ $rc -eq 0 ]; then
  SALT_TC_PAIR=1 timeout 240 python -m pytest tests/test_engine_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_pair_engine_g.log 2>&1
  echo "pytest engine (pair) rc=$?" >> gpurun_out/summary_g.txt
  SALT_TC_PAIR=1 timeout 240 python bench.py --no-extra --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err
  echo "bench pair rc=$?" >> gpurun_out/summary_g.txt
  head -c 400 gpurun_out/bench_pair.json; echo
  SALT_TC_PAIR=1 timeout 120 python profiles/microbench_conv.py > gpurun_out/micro_pair.txt 2>&1
fi
cat gpurun_out/summary_g.txt
