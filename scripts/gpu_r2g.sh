#!/bin/bash
T=${1:-r2g}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/${T}_knobs.log
for pt in 1 0; do for cfg in C3 C5 C2; do
  echo "=== NRB_PASS_THROUGH=$pt $cfg" >> gpurun_out/${T}_knobs.log
  NRB_PASS_THROUGH=$pt python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "frame 7|wave " >> gpurun_out/${T}_knobs.log
done; done
( timeout 2400 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 ) > gpurun_out/${T}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${T}_parity.jsonl 2>/dev/null
( NRB_BUILD_TIMES=1 EXP_BUILDERS=ploc,lbvh,ploc,lbvh timeout 600 python scripts/exp_builders.py C4 ) > gpurun_out/${T}_builders.log 2>&1
grep -E "===|frame 7|wave" gpurun_out/${T}_knobs.log; tail -14 gpurun_out/${T}_pytest.log; grep -E "device builder|create|triangle tree" gpurun_out/${T}_builders.log | tail -24
