#!/bin/bash
T=${1:-r2t7}
mkdir -p gpurun_out
: > gpurun_out/${T}.log
for i in 1 2 3; do NRB_BUILD_TIMES=1 python scripts/bvh_dump.py C4 /tmp/x_C4.bin 2>&1 | grep -E "SAH:|opaque" >> gpurun_out/${T}.log; done
( EXP_BUILDERS=sah,sah timeout 600 python scripts/exp_builders.py C4 C3 ) 2>&1 | grep create >> gpurun_out/${T}.log
cat gpurun_out/${T}.log
