#!/bin/bash
# tail kernel tracing its own light samples (no shadow queue, no second launch): C3 whole frame and 1/8 shard
T=${1:-r2u}
mkdir -p gpurun_out
: > gpurun_out/${T}.log
for v in 0 1; do
  echo "== NRB_TAIL_INLINE_SHADOW=$v C3" >> gpurun_out/${T}.log
  NRB_TAIL_INLINE_SHADOW=$v timeout 300 python scripts/exp_c3.py C3 8 2>&1 | grep -E "frame [5-7]|wave|tail" >> gpurun_out/${T}.log
  echo "== NRB_TAIL_INLINE_SHADOW=$v C3 shard 1/8" >> gpurun_out/${T}.log
  NRB_TAIL_INLINE_SHADOW=$v timeout 300 python scripts/exp_shard.py C3 8 2>&1 | grep -E "frame [3-5]|wave|tail" >> gpurun_out/${T}.log
  echo "== NRB_TAIL_INLINE_SHADOW=$v C2" >> gpurun_out/${T}.log
  NRB_TAIL_INLINE_SHADOW=$v timeout 300 python scripts/exp_c3.py C2 8 2>&1 | grep -E "frame [5-7]" >> gpurun_out/${T}.log
done
NRB_TAIL_INLINE_SHADOW=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -q --timeout 180 2>&1 | tail -3 >> gpurun_out/${T}.log
cat gpurun_out/${T}.log
