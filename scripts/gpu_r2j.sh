#!/bin/bash
T=${1:-r2j}
mkdir -p gpurun_out
rm -f gpurun_out/${T}_shard.log
for world in 8 4; do
  echo "=== shard 1/$world C3" >> gpurun_out/${T}_shard.log
  python scripts/exp_shard.py C3 $world 2>&1 | grep -E "frame [45]|wave|nrb" >> gpurun_out/${T}_shard.log
done
for tr in 65536 262144 1048576; do
  echo "=== shard 1/8 C3 NRB_TAIL_RAYS=$tr" >> gpurun_out/${T}_shard.log
  NRB_TAIL_RAYS=$tr python scripts/exp_shard.py C3 8 2>&1 | grep -E "frame [45]" >> gpurun_out/${T}_shard.log
done
cat gpurun_out/${T}_shard.log
