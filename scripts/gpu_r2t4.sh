#!/bin/bash
T=${1:-r2t4}
mkdir -p gpurun_out
: > gpurun_out/${T}_variants.log
for rep in 1 2; do
for cfg in C3 C4 C5; do
echo "=== head $cfg" >> gpurun_out/${T}_variants.log
timeout 300 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "frame [5-7]" >> gpurun_out/${T}_variants.log
done
bash scripts/run_variants.sh $T "C3 C4 C5" base > /dev/null
done
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -q --timeout 180 2>&1 | tail -2 ) >> gpurun_out/${T}_variants.log
grep -E "===|frame [67]|passed|failed" gpurun_out/${T}_variants.log
