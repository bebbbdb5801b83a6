#!/bin/bash
T=${1:-r2r}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/${T}_pytest.log
for i in 1 2 3; do
( NRB_NODE_FORMAT=2 timeout 200 python -u -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_fuzz_gpu.py -m gpu -v --timeout 60 2>&1 | tail -4 ) >> gpurun_out/${T}_pytest.log
( timeout 200 python -u -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_fullsize_gpu.py -m gpu -v --timeout 60 2>&1 | tail -4 ) >> gpurun_out/${T}_pytest.log
done
for i in 1 2 3; do timeout 120 python scripts/exp_c3.py C4 12 2>&1 | grep -E "frame 11" >> gpurun_out/${T}_pytest.log; done
cat gpurun_out/${T}_pytest.log
