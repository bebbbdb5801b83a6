#!/bin/bash
# GPU call 4: lane-refill trace kernel, tail with shared shadow counter
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c4_pytest.log
for v in "12 20 20" "0 0 0" "0 20 20" "16 24 24" "20 28 28" "8 16 16"; do
  set -- $v
  for cfg in C3 C4; do
    echo "=== refill primary=$1 rays=$2 shadow=$3 $cfg" >> gpurun_out/c4_variants.log
    NRB_REFILL_PRIMARY=$1 NRB_REFILL_RAYS=$2 NRB_REFILL_SHADOW=$3 timeout 300 python scripts/exp_c3.py $cfg 6 >> gpurun_out/c4_variants.log 2>&1
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 6 -c 3 \
     -o gpurun_out/c4_prof_c3 -f python scripts/exp_c3.py C3 3 > gpurun_out/c4_ncu_c3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 6 -c 3 \
     -o gpurun_out/c4_prof_c4 -f python scripts/exp_c3.py C4 3 > gpurun_out/c4_ncu_c4.log 2>&1
tail -3 gpurun_out/c4_pytest.log; grep -E "===|frame 5|wave " gpurun_out/c4_variants.log
