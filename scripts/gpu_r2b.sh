#!/bin/bash
# second GPU call of round 2: FFMA2 micro-benchmark, node-format variants on C3/C4/C2, full GPU test-suite on the default build
T=${1:-r2b}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
./scripts/ubench_ffma2 > gpurun_out/${T}_ubench.txt 2>&1
bash scripts/run_variants.sh $T "C3 C4 C2" f0 f1 f2 > /dev/null 2>&1
( timeout 2400 python -m pytest tests -m gpu -q --durations=15 2>&1 | tail -80 ) > gpurun_out/${T}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${T}_parity.jsonl 2>/dev/null
cat gpurun_out/${T}_ubench.txt; grep -E "===|frame 7" gpurun_out/${T}_variants.log; tail -30 gpurun_out/${T}_pytest.log
