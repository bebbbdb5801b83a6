#!/bin/bash
# Developer tool: build kernel variants side by side (nrays_b200/csrc/variants/lib_<name>.so) for one-call A/B runs on the GPU
# (select with NRB_LIB=...).  usage: bash scripts/build_variants.sh name1="-DFLAG=1 ..." name2="..."
set -e
cd "$(dirname "$0")/../nrays_b200/csrc"
mkdir -p variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  /usr/local/cuda/bin/nvcc $flags -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
     -shared -o variants/lib_${name}.so -x cu kernels.cu api.cu lbvh.cu bvh_build.cpp -lcudart &
done
wait
ls -la variants
