#!/bin/bash
# Round-2 pass O (1 GPU): parity after the fold revert / channel-slab upsample backward, then ncu --set full of the memory-bound
# kernels that still sit far from the HBM roofline (fused upsample backward, the shuffle-based ring ops, fold backward, BN reduce).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" > gpurun_out/r2o_summary.txt
timeout 200 python profiles/step_timeline.py > gpurun_out/r2o_step_timeline.txt 2>&1
echo "timeline rc=$?" >> gpurun_out/r2o_summary.txt
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f"
timeout 300 $NCU -k regex:upsample_bwd_kernel -c 6 -o gpurun_out/r2o_ncu_upsample python profiles/step_timeline.py > gpurun_out/r2o_ncu1.log 2>&1
echo "ncu upsample rc=$?" >> gpurun_out/r2o_summary.txt
timeout 300 $NCU -k regex:"ScseApplyOp|FinalFwdOp" -c 6 -o gpurun_out/r2o_ncu_pixfwd python profiles/step_timeline.py > gpurun_out/r2o_ncu2.log 2>&1
echo "ncu pixel fwd rc=$?" >> gpurun_out/r2o_summary.txt
timeout 300 $NCU -k regex:"ScseBwdOp|fold_bwd_kernel" -c 3 -o gpurun_out/r2o_ncu_pixbwd python profiles/step_timeline.py > gpurun_out/r2o_ncu3.log 2>&1
echo "ncu pixel bwd rc=$?" >> gpurun_out/r2o_summary.txt
timeout 300 $NCU -k regex:"BnBwdReduceOp|BnApplyOp" -c 4 -o gpurun_out/r2o_ncu_bn python profiles/step_timeline.py > gpurun_out/r2o_ncu4.log 2>&1
echo "ncu bn rc=$?" >> gpurun_out/r2o_summary.txt
cat gpurun_out/r2o_summary.txt; tail -3 gpurun_out/r2o_pytest.log; head -1 gpurun_out/r2o_step_timeline.txt
