#!/bin/bash
T=${1:-r2s1}
mkdir -p gpurun_out
: > gpurun_out/${T}_variants.log
for cfg in C3 C5 C2; do
echo "=== main $cfg" >> gpurun_out/${T}_variants.log
timeout 300 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "frame [5-7]" >> gpurun_out/${T}_variants.log
done
bash scripts/run_variants.sh $T "C3 C5 C2" sb64 swarp sb256 > /dev/null
grep -E "===|frame [67]" gpurun_out/${T}_variants.log
