#!/bin/bash
# ncu --set full of the experimental CTA-pair kernel (layer1 / final.0 / layer2 shapes of profiles/microbench_conv.py), for round 2
mkdir -p gpurun_out
SALT_TC_PAIR=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_rows_pair -c 5 -f -o gpurun_out/pair_full_r1i \
  python profiles/microbench_conv.py > gpurun_out/ncu_pair_i.log 2>&1
echo "ncu pair rc=$?"; tail -3 gpurun_out/ncu_pair_i.log
