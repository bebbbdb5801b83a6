#!/bin/bash
T=${1:-r2n}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_knobs.log
run() { echo "=== $1 $2" >> gpurun_out/${T}_knobs.log; env $1 python scripts/exp_c3.py $2 6 2>&1 | grep -E "frame 5" >> gpurun_out/${T}_knobs.log; }
run "NRB_BATCH_SLOTS=8388608" C5
run "NRB_BATCH_SLOTS=16777216 NRB_SHADOW_CAP=16777216" C5
run "NRB_BATCH_SLOTS=33554432 NRB_SHADOW_CAP=33554432" C5
run "NRB_BATCH_SLOTS=67108864 NRB_SHADOW_CAP=67108864" C5
run "NRB_BATCH_SLOTS=33554432 NRB_SHADOW_CAP=33554432" C4
run "NRB_BATCH_SLOTS=33554432 NRB_SHADOW_CAP=33554432" C3
cat gpurun_out/${T}_knobs.log
