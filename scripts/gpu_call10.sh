#!/bin/bash
mkdir -p gpurun_out
for v in base n48; do
  lib=build/variants/lib_$v.so; [ $v = base ] && lib=nrays_b200/csrc/libnrays_b200.so
  for cfg in C3 C4 C2; do
    echo "=== $v $cfg" >> gpurun_out/c10_variants.log
    NRB_LIB=$lib timeout 300 python scripts/exp_c3.py $cfg 8 >> gpurun_out/c10_variants.log 2>&1
  done
done
( NRB_LIB=build/variants/lib_n48.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/c10_pytest_n48.log
NRB_LIB=build/variants/lib_n48.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 6 -c 2 \
     -o gpurun_out/c10_prof_n48 -f python scripts/exp_c3.py C3 3 > gpurun_out/c10_ncu.log 2>&1
tail -3 gpurun_out/c10_pytest_n48.log; grep -E "===|frame [67]|wave " gpurun_out/c10_variants.log
