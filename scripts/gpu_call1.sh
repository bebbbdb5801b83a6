#!/bin/bash
# GPU call 1 of this session: validate the restored tree, compare node-format variants, capture ncu full.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c1_pytest.log
for v in base ch v8 ch_v8; do
  lib=build/variants/lib_$v.so; [ $v = base ] && lib=nrays_b200/csrc/libnrays_b200.so
  for cfg in C3 C4; do
    echo "=== $v $cfg" >> gpurun_out/c1_variants.log
    NRB_LIB=$lib timeout 300 python scripts/exp_c3.py $cfg 6 >> gpurun_out/c1_variants.log 2>&1
  done
done
# parity with the most aggressive variant
( NRB_LIB=build/variants/lib_ch_v8.so timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c1_pytest_chv8.log
for v in base ch_v8; do
  lib=build/variants/lib_$v.so; [ $v = base ] && lib=nrays_b200/csrc/libnrays_b200.so
  NRB_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 8 -c 4 \
     -o gpurun_out/c1_prof_$v -f python scripts/exp_c3.py C3 3 > gpurun_out/c1_ncu_$v.log 2>&1
done
python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -3 gpurun_out/c1_pytest.log; tail -3 gpurun_out/c1_pytest_chv8.log; grep -E "===|frame 5" gpurun_out/c1_variants.log; cat gpurun_out/c1_bench.json | cut -c1-400
