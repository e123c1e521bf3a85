#!/bin/bash
# Round-2 pass E: elementwise passes with more loads in flight - parity (engine + dropin suites), bench, per-kernel launch list with DRAM bytes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_dropin_gpu.py tests/test_baseline_configs_gpu.py tests/test_io_formats.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2e_summary.txt
timeout 300 python bench.py --no-cpu-baseline --no-se50 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?" >> gpurun_out/r2e_summary.txt
SALT_ENGINE_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2e_bench_under_ncu.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/r2e_summary.txt
cat gpurun_out/r2e_summary.txt; tail -3 gpurun_out/r2e_pytest.log; head -c 600 gpurun_out/r2e_bench.json
