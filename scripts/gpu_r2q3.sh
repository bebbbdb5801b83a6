#!/bin/bash
T=${1:-r2q3}
mkdir -p gpurun_out
: > gpurun_out/${T}_variants.log
( timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -m gpu -q --timeout 300 -k "node_format or c4" 2>&1 | tail -5 ) > gpurun_out/${T}_pytest.log
echo "=== main C3" >> gpurun_out/${T}_variants.log
timeout 300 python scripts/exp_c3.py C3 8 2>&1 | grep -E "frame [5-7]" >> gpurun_out/${T}_variants.log
echo "=== main C4" >> gpurun_out/${T}_variants.log
timeout 300 python scripts/exp_c3.py C4 8 2>&1 | grep -E "frame [5-7]" >> gpurun_out/${T}_variants.log
bash scripts/run_variants.sh $T "C3 C4" trinoalloc mb8 mb7 mb10 > /dev/null
cat gpurun_out/${T}_pytest.log; grep -E "===|frame [67]" gpurun_out/${T}_variants.log
