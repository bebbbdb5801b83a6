#!/bin/bash
mkdir -p gpurun_out
for m in 0 2 1; do
  echo "=== inline=$m" >> gpurun_out/c9.log
  NRB_TAIL_INLINE=$m timeout 600 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -k "c1_reference" 2>&1 | tail -4 >> gpurun_out/c9.log
done
NRB_TAIL_INLINE=1 timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "shape_zoo_phong or transparency" 2>&1 | tail -30 >> gpurun_out/c9.log
NRB_TAIL_INLINE=1 timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "transparency" 2>&1 | tail -30 >> gpurun_out/c9.log
cat gpurun_out/c9.log
