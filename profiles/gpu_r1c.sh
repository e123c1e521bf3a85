#!/bin/bash
# Round-1 (session 2) GPU pass 2: parity tests with the cluster/multicast convolution, cluster-size sweep of the bench, launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_c.log 2>&1
echo "pytest(cl4) rc=$?" > gpurun_out/summary_c.txt
tail -4 gpurun_out/pytest_gpu_c.log
if ! grep -q "pytest(cl4) rc=0" gpurun_out/summary_c.txt; then
  SALT_TC_CLUSTER=1 timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_c_cl1.log 2>&1
  echo "pytest(cl1) rc=$?" >> gpurun_out/summary_c.txt
fi
SALT_TC_CLUSTER=2 timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_conv_cl2.log 2>&1
echo "pytest conv (cl2) rc=$?" >> gpurun_out/summary_c.txt
for cl in 4 2 1; do
  SALT_TC_CLUSTER=$cl timeout 300 python bench.py --no-extra --no-cpu-baseline > gpurun_out/bench_cl$cl.json 2> gpurun_out/bench_cl$cl.err
  echo "bench cl$cl rc=$?" >> gpurun_out/summary_c.txt
done
timeout 900 python bench.py > gpurun_out/bench_full_c.json 2> gpurun_out/bench_full_c.err
echo "bench full rc=$?" >> gpurun_out/summary_c.txt
SALT_ENGINE_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8000 --csv \
  --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_under_ncu_c.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/summary_c.txt
cat gpurun_out/summary_c.txt
for f in gpurun_out/bench_cl4.json gpurun_out/bench_cl2.json gpurun_out/bench_cl1.json; do head -c 400 $f; echo; done
