#!/bin/bash
T=${1:-r2t5}
mkdir -p gpurun_out
: > gpurun_out/${T}.log
nproc >> gpurun_out/${T}.log
for d in 8 4 2 1; do for i in 1 2; do echo -n "div $d: " >> gpurun_out/${T}.log; NRB_BVH_TASK_DIV=$d NRB_BUILD_TIMES=1 python scripts/bvh_dump.py C4 /tmp/x_C4.bin 2>&1 | grep -E "SAH:" >> gpurun_out/${T}.log; done; done
cat gpurun_out/${T}.log
