#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/c14_pytest.log
for v in base sq0; do
  lib=build/variants/lib_$v.so; [ $v = base ] && lib=nrays_b200/csrc/libnrays_b200.so
  for w in 8 2; do
    echo "=== $v shard 1/$w" >> gpurun_out/c14.log
    NRB_LIB=$lib python scripts/exp_shard.py C3 $w 2>&1 | grep -E "frame [45]|wave" >> gpurun_out/c14.log
  done
  echo "=== $v full C3" >> gpurun_out/c14.log
  NRB_LIB=$lib python scripts/exp_c3.py C3 6 2>&1 | grep -E "frame [45]|wave" >> gpurun_out/c14.log
done
tail -3 gpurun_out/c14_pytest.log; cat gpurun_out/c14.log
