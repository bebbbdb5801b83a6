#!/bin/bash
T=${1:-r2h}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_knobs.log
for rc in 0 16 24 28 31; do for cfg in C3; do
  echo "=== NRB_REFILL_CHAIN=$rc $cfg" >> gpurun_out/${T}_knobs.log
  NRB_REFILL_CHAIN=$rc python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "frame 7|wave " >> gpurun_out/${T}_knobs.log
done; done
echo "=== NRB_PASS_THROUGH=0 C3" >> gpurun_out/${T}_knobs.log
NRB_PASS_THROUGH=0 python scripts/exp_c3.py C3 8 2>&1 | grep -E "frame 7|wave " >> gpurun_out/${T}_knobs.log
for rc in 24; do for cfg in C5; do
  echo "=== NRB_REFILL_CHAIN=$rc $cfg" >> gpurun_out/${T}_knobs.log
  NRB_REFILL_CHAIN=$rc python scripts/exp_c3.py $cfg 6 2>&1 | grep -E "frame 5" >> gpurun_out/${T}_knobs.log
done; done
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "pass_through or driver_paths or mesh_scene" 2>&1 | tail -5 ) > gpurun_out/${T}_pytest.log
grep -E "===|frame|wave" gpurun_out/${T}_knobs.log; tail -3 gpurun_out/${T}_pytest.log
