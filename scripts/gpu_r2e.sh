#!/bin/bash
T=${1:-r2e}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "device_bvh or alpha_mapped or both_children or exact_ties or rgb8" 2>&1 | tail -30 ) > gpurun_out/${T}_pytest.log
( NRB_BUILD_TIMES=1 timeout 600 python scripts/exp_builders.py C3 C4 ) > gpurun_out/${T}_builders.log 2>&1
for r in 8 32; do ( EXP_BUILDERS=ploc NRB_PLOC_RADIUS=$r timeout 600 python scripts/exp_builders.py C3 C4 ) >> gpurun_out/${T}_builders.log 2>&1; done
tail -12 gpurun_out/${T}_pytest.log; grep -E "create|build:" gpurun_out/${T}_builders.log
