#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c12_pytest.log
for rv in 1 0; do
  for cfg in C3 C4 C2 C1; do
    echo "=== reverse=$rv $cfg" >> gpurun_out/c12_variants.log
    NRB_REVERSE_SHADOW=$rv timeout 300 python scripts/exp_c3.py $cfg 8 >> gpurun_out/c12_variants.log 2>&1
  done
done
tail -3 gpurun_out/c12_pytest.log; grep -E "===|frame [67]|wave " gpurun_out/c12_variants.log
