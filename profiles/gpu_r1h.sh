#!/bin/bash
# confirmation run of the parity suite on the final tree
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_h.log 2>&1
echo "pytest rc=$?" > gpurun_out/summary_h.txt
tail -5 gpurun_out/pytest_gpu_h.log; cat gpurun_out/summary_h.txt
