#!/bin/bash
T=${1:-r2}
mkdir -p gpurun_out
O=gpurun_out/${T}_sanitizer.txt
echo "# compute-sanitizer on scripts/sanitize.py (C1-C4 reduced; node formats 0 / 2 / 3 / 4 incl. the speculative loop and its predicated push / pop; SAH / PLOC / LBVH builders; default / no-tail / tiny-batch+chunked-shadow / dynamic-fetch / forward-shadow paths; float and RGB8 tile exchanges), B200" > $O
for tool in memcheck racecheck synccheck; do
  echo "## $tool" >> $O
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize.py 2>&1 | grep -E "sanitize: done|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|Error" | head -20 >> $O
done
cat $O
