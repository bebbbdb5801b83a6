#!/bin/bash
# GPU call 7: tests; tail threshold / tail occupancy sweeps
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c7_pytest.log
for tr in 4194304 1048576 262144 65536 0; do
  echo "=== tail_rays=$tr C3" >> gpurun_out/c7_variants.log
  NRB_TAIL_RAYS=$tr timeout 300 python scripts/exp_c3.py C3 6 >> gpurun_out/c7_variants.log 2>&1
done
for v in t5 t6; do
  echo "=== $v C3" >> gpurun_out/c7_variants.log
  NRB_LIB=build/variants/lib_$v.so timeout 300 python scripts/exp_c3.py C3 6 >> gpurun_out/c7_variants.log 2>&1
done
tail -3 gpurun_out/c7_pytest.log; grep -E "===|frame [45]|wave " gpurun_out/c7_variants.log
