#!/bin/bash
T=${1:-r2y3}
mkdir -p gpurun_out
( EXP_BUILDERS=sah,lbvh,ploc,ploc timeout 600 python scripts/exp_builders.py C3 C4 ) 2>&1 | grep create > gpurun_out/${T}.log
( timeout 600 python -m pytest tests -m gpu -q --timeout 120 2>&1 | tail -3 ) >> gpurun_out/${T}.log
( NRB_BUILDER=ploc timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py tests/test_fullsize_gpu.py -m gpu -q --timeout 120 2>&1 | tail -3 ) >> gpurun_out/${T}.log
cat gpurun_out/${T}.log
