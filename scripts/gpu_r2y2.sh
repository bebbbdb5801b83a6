#!/bin/bash
T=${1:-r2y2}
mkdir -p gpurun_out
rm -f gpurun_out/${T}.log
for t in 4 8 32; do echo "=== C3 NRB_RESAH_LEAF=$t" >> gpurun_out/${T}.log; ( NRB_RESAH_LEAF=$t EXP_BUILDERS=lbvh,ploc timeout 300 python scripts/exp_builders.py C3 ) 2>&1 | grep create >> gpurun_out/${T}.log; done
for t in 44 88 350; do echo "=== C4 NRB_RESAH_LEAF=$t" >> gpurun_out/${T}.log; ( NRB_RESAH_LEAF=$t EXP_BUILDERS=lbvh,ploc timeout 300 python scripts/exp_builders.py C4 ) 2>&1 | grep create >> gpurun_out/${T}.log; done
cat gpurun_out/${T}.log
