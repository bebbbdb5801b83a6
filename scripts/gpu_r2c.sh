#!/bin/bash
T=${1:-r2c}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/${T}_variants.log
bash scripts/run_variants.sh $T "C3 C4 C2" f0 f1 f2 > /dev/null 2>&1
( timeout 1200 python -m pytest tests/test_parity_gpu.py -m gpu -q --durations=5 2>&1 | tail -30 ) > gpurun_out/${T}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${T}_parity.jsonl 2>/dev/null
grep -E "===|frame 7|wave  [01]" gpurun_out/${T}_variants.log; tail -8 gpurun_out/${T}_pytest.log
