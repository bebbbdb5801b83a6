#!/bin/bash
T=${1:-r2z3}
mkdir -p gpurun_out
( NRB_BUILD_TIMES=1 EXP_BUILDERS=lbvh,lbvh,ploc,ploc timeout 600 python scripts/exp_builders.py C4 ) > gpurun_out/${T}.log 2>&1
cat gpurun_out/${T}.log
