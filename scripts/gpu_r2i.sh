#!/bin/bash
T=${1:-r2i}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/${T}_knobs.log
for u in 1 0; do for cfg in C3 C5; do
  echo "=== NRB_UNIFIED_TREE=$u $cfg" >> gpurun_out/${T}_knobs.log
  NRB_UNIFIED_TREE=$u python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "frame 7|wave  [0-3]" >> gpurun_out/${T}_knobs.log
done; done
( timeout 2400 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -30 ) > gpurun_out/${T}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${T}_parity.jsonl 2>/dev/null
grep -E "===|frame 7|wave" gpurun_out/${T}_knobs.log; tail -12 gpurun_out/${T}_pytest.log
