#!/bin/bash
T=${1:-r2q}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 400 python -u -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_fullsize_gpu.py -m gpu -v --timeout 60 -x 2>&1 | tail -80 ) > gpurun_out/${T}_pytest.log
grep -E "PASSED|FAILED|Timeout|ERROR" gpurun_out/${T}_pytest.log | tail -30
