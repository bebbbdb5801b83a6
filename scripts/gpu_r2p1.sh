#!/bin/bash
# pass-through of fully transparent texels inside the traversal: parity + timing (NRB_PASS_THROUGH=0 = chains through the queues)
T=${1:-r2p1}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
( timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 ) > gpurun_out/${T}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${T}_parity.jsonl 2>/dev/null
: > gpurun_out/${T}.log
for cfg in C3 C5; do
  for v in 0 1; do
    echo "== $cfg NRB_PASS_THROUGH=$v" >> gpurun_out/${T}.log
    NRB_PASS_THROUGH=$v timeout 300 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "^frame [5-7]|wave " >> gpurun_out/${T}.log
  done
done
echo "== C3 shard 1/8" >> gpurun_out/${T}.log
timeout 300 python scripts/exp_shard.py C3 8 2>&1 | grep -E "frame [3-5]|wave" >> gpurun_out/${T}.log
cat gpurun_out/${T}_pytest.log gpurun_out/${T}.log
