#!/bin/bash
T=${1:-r2t3}
mkdir -p gpurun_out
: > gpurun_out/${T}.log
for cfg in C4 C3 C5; do
  for c in none 0 6 25; do
    echo "== $cfg NRB_CARVEOUT=$c" >> gpurun_out/${T}.log
    if [ $c = none ]; then timeout 300 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "^frame [6-7]" >> gpurun_out/${T}.log
    else NRB_CARVEOUT=$c timeout 300 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "^frame [6-7]" >> gpurun_out/${T}.log; fi
  done
done
cat gpurun_out/${T}.log
