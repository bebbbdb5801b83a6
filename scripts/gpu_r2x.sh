#!/bin/bash
T=${1:-r2x}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_knobs.log
run() { echo "=== $1 $2" >> gpurun_out/${T}_knobs.log; env $1 timeout 120 python scripts/exp_c3.py $2 8 2>&1 | grep -E "frame 7|wave  [01]" >> gpurun_out/${T}_knobs.log; }
run "NRB_X=0" C3
run "NRB_X=0" C2
run "NRB_X=0" C4
run "NRB_X=0" C5
( timeout 900 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -5 ) > gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_knobs.log; cat gpurun_out/${T}_pytest.log
