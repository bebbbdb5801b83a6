#!/bin/bash
# GPU call 5: dynamic fetch with patience
mkdir -p gpurun_out
for v in "0 0 0 1" "12 20 20 4" "12 20 20 8" "12 20 20 16" "16 24 24 8" "0 20 20 8" "12 16 16 32"; do
  set -- $v
  for cfg in C3 C4; do
    echo "=== refill primary=$1 rays=$2 shadow=$3 patience=$4 $cfg" >> gpurun_out/c5_variants.log
    NRB_REFILL_PRIMARY=$1 NRB_REFILL_RAYS=$2 NRB_REFILL_SHADOW=$3 NRB_REFILL_PATIENCE=$4 timeout 300 python scripts/exp_c3.py $cfg 6 >> gpurun_out/c5_variants.log 2>&1
  done
done
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c5_pytest.log
tail -3 gpurun_out/c5_pytest.log; grep -E "===|frame [345]" gpurun_out/c5_variants.log
