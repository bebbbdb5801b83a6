#!/bin/bash
# L2 access-policy window over the node array: C4, C5, C3
T=${1:-r2z1}
mkdir -p gpurun_out
: > gpurun_out/${T}.log
for cfg in C4 C5 C3; do
  for mb in 0 16 64 128; do
    echo "== $cfg NRB_L2_PERSIST_MB=$mb" >> gpurun_out/${T}.log
    NRB_BUILD_TIMES=1 NRB_L2_PERSIST_MB=$mb timeout 300 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "^frame [5-7]|L2 persistence" >> gpurun_out/${T}.log
  done
done
cat gpurun_out/${T}.log
