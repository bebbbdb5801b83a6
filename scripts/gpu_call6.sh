#!/bin/bash
# GPU call 6: fastdiv / cheaper shade math; full test + bench
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/c6_pytest.log
for cfg in C3 C4 C2 C1; do
  echo "=== $cfg" >> gpurun_out/c6_variants.log
  timeout 300 python scripts/exp_c3.py $cfg 6 >> gpurun_out/c6_variants.log 2>&1
done
python bench.py --steps 10 --warmup 3 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
tail -3 gpurun_out/c6_pytest.log; grep -E "===|frame [45]|wave " gpurun_out/c6_variants.log; cut -c1-300 gpurun_out/c6_bench.json
