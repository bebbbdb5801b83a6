#!/bin/bash
T=${1:-r2t}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_knobs.log
run() { echo "=== $1 $2" >> gpurun_out/${T}_knobs.log; env $1 timeout 120 python scripts/exp_c3.py $2 6 2>&1 | grep -E "frame 5" >> gpurun_out/${T}_knobs.log; }
for rr in 0 12 16 24 28; do run "NRB_REFILL_RAYS=$rr NRB_REFILL_SHADOW=$rr" C4; done
for rp in 12 20 28; do run "NRB_REFILL_PRIMARY=$rp" C4; done
run "NRB_REVERSE_SHADOW=0" C4
run "NRB_SMALL_QUEUE=0" C4
cat gpurun_out/${T}_knobs.log
