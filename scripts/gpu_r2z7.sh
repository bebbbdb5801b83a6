#!/bin/bash
T=${1:-r2z7}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 300 -k "node_format or grid_format" 2>&1 | tail -5 ) > gpurun_out/${T}.log
( NRB_BUILD_TIMES=1 timeout 300 python scripts/exp_c3.py C4 4 2>&1 | grep -E "grid|frame 3" ) >> gpurun_out/${T}.log
cat gpurun_out/${T}.log
