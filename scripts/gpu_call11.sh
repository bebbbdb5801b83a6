#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c11_pytest.log
for v in base s7 s8; do
  lib=build/variants/lib_$v.so; [ $v = base ] && lib=nrays_b200/csrc/libnrays_b200.so
  for cfg in C3 C2; do
    echo "=== $v $cfg" >> gpurun_out/c11_variants.log
    NRB_LIB=$lib timeout 300 python scripts/exp_c3.py $cfg 8 >> gpurun_out/c11_variants.log 2>&1
  done
done
for cfg in C4 C1; do
    echo "=== base $cfg" >> gpurun_out/c11_variants.log
    timeout 300 python scripts/exp_c3.py $cfg 8 >> gpurun_out/c11_variants.log 2>&1
done
tail -3 gpurun_out/c11_pytest.log; grep -E "===|frame [67]|wave " gpurun_out/c11_variants.log
