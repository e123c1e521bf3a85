#!/bin/bash
# Round-2 pass W (1 GPU), the record pass: full GPU suite, smoke, the bench line (+ reference arm), the ncu launch list with DRAM
# bytes, ncu --set full of the dominant kernels, racecheck of one small step.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2w_summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2w_summary.txt
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
echo "bench rc=$?" >> gpurun_out/r2w_summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err
echo "bench reference rc=$?" >> gpurun_out/r2w_summary.txt
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/r2w_launches.csv python profiles/one_step.py 2 > gpurun_out/r2w_one_step_under_ncu.log 2>&1
echo "ncu list rc=$?" >> gpurun_out/r2w_summary.txt
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 400 $NCU -k regex:conv_tc_rows -c 6 -o gpurun_out/r2w_ncu_rows python profiles/microbench_conv.py > gpurun_out/r2w_ncu_rows.log 2>&1
echo "ncu rows rc=$?" >> gpurun_out/r2w_summary.txt
timeout 400 $NCU -k regex:conv_wgrad_halo -c 6 -o gpurun_out/r2w_ncu_wgrad python profiles/one_step.py 1 > gpurun_out/r2w_ncu_wgrad.log 2>&1
echo "ncu wgrad rc=$?" >> gpurun_out/r2w_summary.txt
timeout 400 $NCU -k regex:"BnBwdApplyOp|BnApplyOp|BnBwdReduceOp" -c 6 -o gpurun_out/r2w_ncu_ring python profiles/one_step.py 1 > gpurun_out/r2w_ncu_ring.log 2>&1
echo "ncu ring rc=$?" >> gpurun_out/r2w_summary.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python profiles/sanitize_step.py > gpurun_out/r2w_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2w_summary.txt
cat gpurun_out/r2w_summary.txt; tail -3 gpurun_out/r2w_pytest.log; tail -4 gpurun_out/r2w_smoke.log; head -c 400 gpurun_out/r2w_bench.json; echo; head -c 300 gpurun_out/r2w_bench_reference.json; echo; tail -3 gpurun_out/r2w_racecheck.log
