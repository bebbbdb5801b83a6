#!/bin/bash
T=${1:-r2u}
mkdir -p gpurun_out
( NRB_BUILD_TIMES=1 EXP_BUILDERS=sah,lbvh,ploc,lbvh,ploc,sah timeout 600 python scripts/exp_builders.py C4 ) > gpurun_out/${T}_builders.log 2>&1
( EXP_BUILDERS=sah,lbvh,ploc timeout 600 python scripts/exp_builders.py C3 ) >> gpurun_out/${T}_builders.log 2>&1
( timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "device_bvh" 2>&1 | tail -3 ) >> gpurun_out/${T}_builders.log
grep -E "create|device builder|build:|passed|failed" gpurun_out/${T}_builders.log
