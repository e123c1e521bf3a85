#!/bin/bash
# Round-1 final GPU pass: parity tests (default config: 2-CTA clusters with multicast weights in the row-halo conv, CUDA graphs),
# the cluster variant of the tap-table kernel (env), smoke, bench (both arms), launch list with DRAM bytes, ncu --set full.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_e.log 2>&1
echo "pytest rc=$?" > gpurun_out/summary_e.txt
tail -3 gpurun_out/pytest_gpu_e.log
SALT_TC_CLUSTER_GENERIC=2 timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_engine_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_generic_cl2_e.log 2>&1
echo "pytest conv+engine (generic cl2) rc=$?" >> gpurun_out/summary_e.txt
SALT_TC_CLUSTER_GENERIC=4 timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_generic_cl4_e.log 2>&1
echo "pytest conv (generic cl4) rc=$?" >> gpurun_out/summary_e.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_e.log 2>&1
echo "smoke rc=$?" >> gpurun_out/summary_e.txt
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?" >> gpurun_out/summary_e.txt
SALT_TC_CLUSTER_GENERIC=2 timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/bench_generic_cl2.json 2> gpurun_out/bench_generic_cl2.err
echo "bench generic cl2 rc=$?" >> gpurun_out/summary_e.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference rc=$?" >> gpurun_out/summary_e.txt
SALT_ENGINE_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8000 --csv \
  --log-file gpurun_out/launches_r1e.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_under_ncu_e.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/summary_e.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_rows -c 6 -f -o gpurun_out/rows_full_r1e \
  python profiles/microbench_conv.py > gpurun_out/ncu_full_e.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/summary_e.txt
cat gpurun_out/summary_e.txt; head -c 700 gpurun_out/bench_final.json; echo; head -c 300 gpurun_out/bench_generic_cl2.json; echo; cat gpurun_out/bench_reference.json
