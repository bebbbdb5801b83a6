#!/bin/bash
T=${1:-r2t1}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_fuzz_gpu.py -m gpu -q --timeout 600 -s -k "mesh_scenes" 2>&1 | tail -12 ) > gpurun_out/${T}.log
cat gpurun_out/${T}.log
