#!/bin/bash
# Round-2 pass AB (1 GPU): UNetSeResNetXt (grouped convolutions as block-diagonal dense) - golden fixture fp32 (tensor-core and FMA
# forward) and the bf16 mode test; plus the engine suite as a regression check of the shared code paths.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py -m gpu -q -p no:cacheprovider -s -k "sex50 or UNetSeResNetXt" > gpurun_out/r2ab_pytest_sex.log 2>&1
echo "pytest sex rc=$?" > gpurun_out/r2ab_summary.txt
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_baseline_configs_gpu.py tests/test_dropin_gpu.py -m gpu -q -p no:cacheprovider -k "not sex50 and not UNetSeResNetXt" > gpurun_out/r2ab_pytest_rest.log 2>&1
echo "pytest rest rc=$?" >> gpurun_out/r2ab_summary.txt
cat gpurun_out/r2ab_summary.txt; grep -E "passed|failed|golden (eval|train) logits|bf16 (eval|train) logits|Error" gpurun_out/r2ab_pytest_sex.log | head -20; tail -3 gpurun_out/r2ab_pytest_rest.log
