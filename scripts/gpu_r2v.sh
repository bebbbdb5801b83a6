#!/bin/bash
T=${1:-r2v}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_knobs.log
for mw in 2 1; do
  echo "=== NRB_TAIL_MIN_WAVE=$mw C3 full" >> gpurun_out/${T}_knobs.log
  NRB_TAIL_MIN_WAVE=$mw timeout 120 python scripts/exp_c3.py C3 6 2>&1 | grep -E "frame 5|wave" >> gpurun_out/${T}_knobs.log
  echo "=== NRB_TAIL_MIN_WAVE=$mw C3 shard 1/8" >> gpurun_out/${T}_knobs.log
  NRB_TAIL_MIN_WAVE=$mw timeout 120 python scripts/exp_shard.py C3 8 2>&1 | grep -E "frame 5|wave" >> gpurun_out/${T}_knobs.log
done
cat gpurun_out/${T}_knobs.log
