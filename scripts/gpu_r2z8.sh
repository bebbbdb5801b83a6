#!/bin/bash
T=${1:-r2z8}
mkdir -p gpurun_out
( NRB_BUILD_TIMES=1 EXP_BUILDERS=sah,sah,sah timeout 600 python scripts/exp_builders.py C4 C3 ) > gpurun_out/${T}.log 2>&1
grep -E "opaque|unified|create" gpurun_out/${T}.log
