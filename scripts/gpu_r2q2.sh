#!/bin/bash
T=${1:-r2q2}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 180 -k "node_format" 2>&1 | tail -5 ) > gpurun_out/${T}_pytest.log
: > gpurun_out/${T}.log
for cfg in C3 C4 C5; do
  for f in 0 5 2; do
    echo "== $cfg NRB_NODE_FORMAT=$f" >> gpurun_out/${T}.log
    ( NRB_NODE_FORMAT=$f timeout 300 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "^frame [5-7]|error|Error" | grep -v "^  " ) >> gpurun_out/${T}.log
  done
done
cat gpurun_out/${T}_pytest.log gpurun_out/${T}.log
