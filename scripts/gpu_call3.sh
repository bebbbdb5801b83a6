#!/bin/bash
# GPU call 3: ray pairs + merged tail
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c3_pytest.log
for pw in 2 1 0 3; do
  for cfg in C3 C4; do
    echo "=== pair_waves=$pw $cfg" >> gpurun_out/c3_variants.log
    NRB_PAIR_WAVES=$pw timeout 300 python scripts/exp_c3.py $cfg 6 >> gpurun_out/c3_variants.log 2>&1
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 8 -c 3 \
     -o gpurun_out/c3_prof_pairs -f python scripts/exp_c3.py C3 3 > gpurun_out/c3_ncu.log 2>&1
tail -3 gpurun_out/c3_pytest.log; grep -E "===|frame 5|wave " gpurun_out/c3_variants.log
