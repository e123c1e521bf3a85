#!/bin/bash
# Round-1 (session 2) GPU pass: parity tests, bench (CUDA graphs on/off, old/new narrow-layer epilogue), ncu launch list with DRAM
# bytes, one `ncu --set full` capture of the row-halo convolution kernel.   usage: gpurun -- bash profiles/gpu_r1b.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" > gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err
echo "bench_graph rc=$?" >> gpurun_out/summary.txt
SALT_ENGINE_GRAPH=0 timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err
echo "bench_eager rc=$?" >> gpurun_out/summary.txt
SALT_ENGINE_GRAPH=0 SALT_TC_DEBUG=32 timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/bench_eager_oldepi.json 2> gpurun_out/bench_eager_oldepi.err
timeout 200 python profiles/microbench_conv.py > gpurun_out/micro_new.txt 2>&1
SALT_TC_DEBUG=32 timeout 200 python profiles/microbench_conv.py > gpurun_out/micro_oldepi.txt 2>&1
SALT_ENGINE_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8000 --csv \
  --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_rows -c 6 -f -o gpurun_out/rows_full_r1b \
  python profiles/microbench_conv.py > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; head -c 1500 gpurun_out/bench_graph.json; echo; head -c 600 gpurun_out/bench_eager.json; echo; head -c 600 gpurun_out/bench_eager_oldepi.json
