#!/bin/bash
# the full randomised parity sweep on the final tree (three seeds)
T=${1:-r2fz}
mkdir -p gpurun_out
: > gpurun_out/${T}.log
for spec in "150 1" "200 7" "200 23"; do
  echo "## python scripts/fuzz_parity.py $spec" >> gpurun_out/${T}.log
  timeout 900 python scripts/fuzz_parity.py $spec 2>&1 | grep -E "BAD|chaotic|^fuzz:" >> gpurun_out/${T}.log
done
cat gpurun_out/${T}.log | cut -c1-400
