#!/bin/bash
T=${1:-r2f}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
( NRB_BUILD_TIMES=1 timeout 600 python scripts/exp_builders.py C3 C4 ) > gpurun_out/${T}_builders.log 2>&1
for top in 0 65536; do for cfg in C3 C4; do
  echo "=== NRB_RELAYOUT_TOP=$top $cfg" >> gpurun_out/${T}_knobs.log
  NRB_RELAYOUT_TOP=$top python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "frame 7" >> gpurun_out/${T}_knobs.log
done; done
( timeout 2400 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 ) > gpurun_out/${T}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${T}_parity.jsonl 2>/dev/null
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
grep -E "create" gpurun_out/${T}_builders.log; grep -E "===|frame 7" gpurun_out/${T}_knobs.log; tail -14 gpurun_out/${T}_pytest.log; cut -c1-300 gpurun_out/${T}_bench.json
