#!/bin/bash
T=${1:-r2d}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_variants.log gpurun_out/${T}_knobs.log
bash scripts/run_variants.sh $T "C3" base t5 t6 s7 s8 > /dev/null 2>&1
for tr in 4194304 1048576 500000 250000 100000 0; do
  echo "=== NRB_TAIL_RAYS=$tr C3" >> gpurun_out/${T}_knobs.log
  NRB_TAIL_RAYS=$tr python scripts/exp_c3.py C3 8 2>&1 | grep -E "frame 7|wave " >> gpurun_out/${T}_knobs.log
done
for f in 0 2; do
  for cfg in C3 C4; do
  echo "=== NRB_NODE_FORMAT=$f $cfg" >> gpurun_out/${T}_knobs.log
  NRB_NODE_FORMAT=$f python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "frame 7" >> gpurun_out/${T}_knobs.log
  done
done
( timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q --durations=5 2>&1 | tail -30 ) > gpurun_out/${T}_pytest.log
grep -E "===|frame 7" gpurun_out/${T}_variants.log gpurun_out/${T}_knobs.log; tail -8 gpurun_out/${T}_pytest.log
