#!/bin/bash
T=${1:-r2t2}
mkdir -p gpurun_out
: > gpurun_out/${T}.log
for f in 4 2 4 2; do
  echo "== C4 NRB_NODE_FORMAT=$f" >> gpurun_out/${T}.log
  NRB_NODE_FORMAT=$f timeout 300 python scripts/exp_c3.py C4 8 2>&1 | grep -E "^frame [5-7]" >> gpurun_out/${T}.log
done
cat gpurun_out/${T}.log
