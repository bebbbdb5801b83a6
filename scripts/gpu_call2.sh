#!/bin/bash
# GPU call 2: merged tail + node load patterns; ncu of C4's trace kernels.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c2_pytest.log
for v in base l0 l2 b8 l2b8; do
  lib=build/variants/lib_$v.so; [ $v = base ] && lib=nrays_b200/csrc/libnrays_b200.so
  for cfg in C3 C4 C2; do
    echo "=== $v $cfg" >> gpurun_out/c2_variants.log
    NRB_LIB=$lib timeout 300 python scripts/exp_c3.py $cfg 6 >> gpurun_out/c2_variants.log 2>&1
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 6 -c 3 \
     -o gpurun_out/c2_prof_c4 -f python scripts/exp_c3.py C4 3 > gpurun_out/c2_ncu_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tail_kernel|shade_kernel" -s 4 -c 3 \
     -o gpurun_out/c2_prof_c3_tail -f python scripts/exp_c3.py C3 3 > gpurun_out/c2_ncu_c3_tail.log 2>&1
tail -3 gpurun_out/c2_pytest.log; grep -E "===|frame 5|wave " gpurun_out/c2_variants.log
