#!/bin/bash
T=${1:-r2z2}
mkdir -p gpurun_out
: > gpurun_out/${T}_variants.log
for cfg in C3 C4 C5; do
echo "=== main $cfg" >> gpurun_out/${T}_variants.log
timeout 300 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "frame [5-7]" >> gpurun_out/${T}_variants.log
done
bash scripts/run_variants.sh $T "C3 C4 C5" vptx > /dev/null
NRB_LIB=$PWD/nrays_b200/csrc/variants/lib_vptx.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -q --timeout 180 2>&1 | tail -3 >> gpurun_out/${T}_variants.log
grep -E "===|frame [67]|passed|failed" gpurun_out/${T}_variants.log
