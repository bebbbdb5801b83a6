#!/bin/bash
# node formats 0 / 2 / 3 / 4: parity tests, then frame times on C3 / C4 / C5
T=${1:-r2q}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 180 -k "node_format or driver_paths" 2>&1 | tail -15 ) > gpurun_out/${T}_pytest.log
: > gpurun_out/${T}.log
for cfg in C3 C4 C5; do
  for f in 0 2 3 4; do
    echo "== $cfg NRB_NODE_FORMAT=$f" >> gpurun_out/${T}.log
    ( NRB_NODE_FORMAT=$f timeout 300 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "^frame [4-7]|error|Error" | grep -v "^  " ) >> gpurun_out/${T}.log
  done
done
cat gpurun_out/${T}_pytest.log gpurun_out/${T}.log
