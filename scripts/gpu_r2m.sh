#!/bin/bash
T=${1:-r2m}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_knobs.log
run() { echo "=== $1 $2" >> gpurun_out/${T}_knobs.log; env $1 python scripts/exp_c3.py $2 8 2>&1 | grep -E "frame 7" >> gpurun_out/${T}_knobs.log; }
for bs in 1048576 2097152 4194304 8388608; do run "NRB_BATCH_SLOTS=$bs" C3; done
for rp in 0 12 20 26; do run "NRB_REFILL_PRIMARY=$rp" C4; done
for rr in 12 16 24 28; do run "NRB_REFILL_RAYS=$rr NRB_REFILL_SHADOW=$rr" C4; done
run "NRB_REVERSE_SHADOW=0" C4
run "NRB_SMALL_QUEUE=0" C4
run "NRB_BATCH_SLOTS=16777216" C4
run "NRB_BATCH_SLOTS=4194304" C4
cat gpurun_out/${T}_knobs.log
