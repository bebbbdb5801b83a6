#!/bin/bash
# last validation of the tree: full -m gpu suite, smoke(), default bench run (wall time), reference arm
T=${1:-r2g}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
( timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > gpurun_out/${T}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${T}_parity.jsonl 2>/dev/null
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/${T}_smoke.log 2>&1
( time python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err ) 2> gpurun_out/${T}_bench_time.log
cat gpurun_out/${T}_pytest.log; tail -5 gpurun_out/${T}_smoke.log; cat gpurun_out/${T}_bench_time.log; cut -c1-300 gpurun_out/${T}_bench.json
