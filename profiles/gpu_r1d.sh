#!/bin/bash
# GPU pass 3: the thread-block-cluster (TMA-multicast) row-halo convolution - parity, then cluster-size sweep of the bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_d.log 2>&1
echo "pytest(cl4) rc=$?" > gpurun_out/summary_d.txt
tail -4 gpurun_out/pytest_gpu_d.log
SALT_TC_CLUSTER=2 timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_engine_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_cl2_d.log 2>&1
echo "pytest conv+engine (cl2) rc=$?" >> gpurun_out/summary_d.txt
for cl in 4 2 1; do
  SALT_TC_CLUSTER=$cl timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/bench_d_cl$cl.json 2> gpurun_out/bench_d_cl$cl.err
  echo "bench cl$cl rc=$?" >> gpurun_out/summary_d.txt
done
SALT_TC_CLUSTER=4 timeout 200 python profiles/microbench_conv.py > gpurun_out/micro_d_cl4.txt 2>&1
SALT_ENGINE_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8000 --csv \
  --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_under_ncu_d.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/summary_d.txt
cat gpurun_out/summary_d.txt
for f in gpurun_out/bench_d_cl4.json gpurun_out/bench_d_cl2.json gpurun_out/bench_d_cl1.json; do head -c 300 $f; echo; done
