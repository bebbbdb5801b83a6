#!/bin/bash
T=${1:-r2y}
mkdir -p gpurun_out
g++ -O2 -std=c++17 -o /tmp/bvh_sim scripts/bvh_sim.cpp
rm -f gpurun_out/${T}_sim.log
for b in lbvh ploc; do
  for t in 0 default 64 1024; do
  echo "=== $b C3 resah-leaf $t" >> gpurun_out/${T}_sim.log
  if [ $t = 0 ]; then export NRB_RESAH_TOP=0; unset NRB_RESAH_LEAF; elif [ $t = default ]; then unset NRB_RESAH_TOP; unset NRB_RESAH_LEAF; else unset NRB_RESAH_TOP; export NRB_RESAH_LEAF=$t; fi
  NRB_CHECK_BVH=1 NRB_BUILDER=$b NRB_DUMP_BVH=/tmp/c3_$b.bin python -c "
import sys; sys.path.insert(0,'.')
from nrays_b200 import configs
s,c,cfg=configs.build('C3'); print('nodes', s.build_info().bvh_nodes, 'depth', s.build_info().max_depth, 'ms', s.build_info().build_ms); s.close()" >> gpurun_out/${T}_sim.log 2>&1
  /tmp/bvh_sim /tmp/c3_$b.bin -250 50 0 0 50 0 45 1920 1080 4 | tail -1 >> gpurun_out/${T}_sim.log
  done
done
unset NRB_RESAH_TOP; unset NRB_RESAH_LEAF
( EXP_BUILDERS=sah,lbvh,ploc timeout 600 python scripts/exp_builders.py C3 C4 ) >> gpurun_out/${T}_sim.log 2>&1
( timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "device_bvh" 2>&1 | tail -3 ) >> gpurun_out/${T}_sim.log
cat gpurun_out/${T}_sim.log
