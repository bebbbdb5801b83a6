#!/bin/bash
T=${1:-r2z4}
mkdir -p gpurun_out
( NRB_BUILD_TIMES=1 EXP_BUILDERS=lbvh,lbvh,ploc,ploc timeout 600 python scripts/exp_builders.py C4 ) > gpurun_out/${T}.log 2>&1
( EXP_BUILDERS=sah,lbvh,ploc,ploc timeout 600 python scripts/exp_builders.py C3 ) 2>&1 | grep create >> gpurun_out/${T}.log
( NRB_CHECK_BVH=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 180 -k "device_bvh or builder or lbvh or ploc" 2>&1 | tail -3 ) >> gpurun_out/${T}.log
( NRB_BUILDER=ploc NRB_CHECK_BVH=1 timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_golden.py -m gpu -q --timeout 180 2>&1 | tail -3 ) >> gpurun_out/${T}.log
grep -v "device builder:" gpurun_out/${T}.log
