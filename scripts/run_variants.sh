#!/bin/bash
# usage: bash scripts/run_variants.sh <tag> "<configs>" name1 name2 ...   (runs scripts/exp_c3.py per variant and config)
T=$1; CFGS=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  for cfg in $CFGS; do
    echo "=== $v $cfg" >> gpurun_out/${T}_variants.log
    NRB_LIB=$PWD/nrays_b200/csrc/variants/lib_${v}.so timeout 300 python scripts/exp_c3.py $cfg 8 2>&1 | grep -E "frame [4-7]|wave " >> gpurun_out/${T}_variants.log
  done
done
grep -E "===|frame 7" gpurun_out/${T}_variants.log
