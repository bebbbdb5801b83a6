#!/bin/bash
# GPU call 8: inline tail
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c8_pytest.log
for v in "1 4194304" "0 4194304" "1 1048576" "1 16777216" "1 262144"; do
  set -- $v
  for cfg in C3; do
    echo "=== inline=$1 tail_rays=$2 $cfg" >> gpurun_out/c8_variants.log
    NRB_TAIL_INLINE=$1 NRB_TAIL_RAYS=$2 timeout 300 python scripts/exp_c3.py $cfg 6 >> gpurun_out/c8_variants.log 2>&1
  done
done
for cfg in C4 C2 C1; do
    echo "=== default $cfg" >> gpurun_out/c8_variants.log
    timeout 300 python scripts/exp_c3.py $cfg 6 >> gpurun_out/c8_variants.log 2>&1
done
tail -3 gpurun_out/c8_pytest.log; grep -E "===|frame [45]|wave " gpurun_out/c8_variants.log
