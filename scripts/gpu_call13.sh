#!/bin/bash
# sanitizer on the new kernels
mkdir -p gpurun_out
{ echo "# compute-sanitizer on scripts/sanitize.py (C1-C4 reduced; default / no-tail / tiny-batch+chunked-shadow / dynamic-fetch / forward-shadow paths), B200";
  echo "## memcheck"; timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize.py 2>&1 | grep -E "sanitize: done|ERROR SUMMARY|Invalid|Error" | head -20;
  echo "## racecheck"; timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize.py 2>&1 | grep -E "sanitize: done|RACECHECK SUMMARY|hazard" | head -20;
  echo "## synccheck"; timeout 900 compute-sanitizer --tool synccheck python scripts/sanitize.py 2>&1 | grep -E "sanitize: done|ERROR SUMMARY|Barrier|divergent" | head -20; } > gpurun_out/c13_sanitizer.txt
cat gpurun_out/c13_sanitizer.txt
