#!/bin/bash
T=${1:-r2p}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_knobs.log
run() { echo "=== $1 $2" >> gpurun_out/${T}_knobs.log; env $1 python scripts/exp_c3.py $2 6 2>&1 | grep -E "frame 5|wave  [01]" >> gpurun_out/${T}_knobs.log; }
run "NRB_X=0" C4
run "NRB_X=0" C3
run "NRB_NODE_FORMAT=2" C3
run "NRB_NODE_FORMAT=0" C4
( NRB_NODE_FORMAT=2 timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_fuzz_gpu.py -m gpu -q 2>&1 | tail -3 ) > gpurun_out/${T}_pytest.log
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_fullsize_gpu.py -m gpu -q 2>&1 | tail -3 ) >> gpurun_out/${T}_pytest.log
cat gpurun_out/${T}_knobs.log; cat gpurun_out/${T}_pytest.log
