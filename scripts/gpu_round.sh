#!/bin/bash
# One GPU call: GPU test-suite, bench line (ours + reference arm), ncu launch list, ncu --set full of one C3 frame.
# usage: bash scripts/gpu_round.sh <tag>      (outputs: gpurun_out/<tag>_*)
T=${1:-rX}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/${T}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trace_kernel|shade_kernel|tail_kernel" -s 7 -c 7 \
    -o gpurun_out/${T}_prof -f python scripts/exp_c3.py C3 3 > gpurun_out/${T}_ncu.log 2>&1
for cfg in C3 C4 C2 C1; do
  echo "=== $cfg" >> gpurun_out/${T}_configs.log
  timeout 300 python scripts/exp_c3.py $cfg 6 >> gpurun_out/${T}_configs.log 2>&1
done
tail -3 gpurun_out/${T}_pytest.log; cut -c1-400 gpurun_out/${T}_bench.json; cut -c1-300 gpurun_out/${T}_bench_ref.json; grep -E "===|frame 5|wave " gpurun_out/${T}_configs.log
