#!/bin/bash
T=${1:-r2l}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_shard.log
for world in 8 1; do
  echo "=== shard 1/$world C3" >> gpurun_out/${T}_shard.log
  python scripts/exp_shard.py C3 $world 2>&1 | grep -E "frame [45]|wave|nrb" >> gpurun_out/${T}_shard.log
done
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -q 2>&1 | tail -4 ) > gpurun_out/${T}_pytest.log
python bench.py --steps 10 --warmup 3 --extra C5 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_shard.log; tail -3 gpurun_out/${T}_pytest.log; cut -c1-260 gpurun_out/${T}_bench.json
