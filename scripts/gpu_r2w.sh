#!/bin/bash
T=${1:-r2w}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_knobs.log
run() { echo "=== $1 $2" >> gpurun_out/${T}_knobs.log; env $1 timeout 120 python scripts/exp_c3.py $2 6 2>&1 | grep -E "frame 5" >> gpurun_out/${T}_knobs.log; }
run "NRB_BVH_CNODE=1.2" C4
run "NRB_BVH_CNODE=2.0" C4
run "NRB_BVH_CNODE=3.0" C4
run "NRB_BVH_CNODE=0.6" C4
run "NRB_BVH_LEAF=2" C4
run "NRB_BVH_CNODE=2.0" C3
run "NRB_BVH_CNODE=0.6" C3
cat gpurun_out/${T}_knobs.log
