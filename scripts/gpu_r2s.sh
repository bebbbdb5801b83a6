#!/bin/bash
T=${1:-r2s2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/${T}_pytest.log
( NRB_NODE_FORMAT=2 timeout 300 python -u -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_fuzz_gpu.py tests/test_bruteforce_pin.py -m gpu -q --timeout 90 2>&1 | tail -6 ) >> gpurun_out/${T}_pytest.log
( timeout 600 python -u -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -6 ) >> gpurun_out/${T}_pytest.log
for cfg in C4 C3 C5; do timeout 120 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "frame 7" >> gpurun_out/${T}_pytest.log; done
cat gpurun_out/${T}_pytest.log
