#!/bin/bash
# developer tool: print the node-visit loop of one trace_kernel instantiation, e.g. scripts/sass_loop.sh ILb0ELb1ELi3 [lines]
cd "$(dirname "$0")/.."
pat="trace_kernel$1"
cuobjdump -sass nrays_b200/csrc/libnrays_b200.so | awk -v p="$pat" '/Function : /{f=index($0,p)>0} f' \
  | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##' > /tmp/sass_loop.txt
L=$(grep -n "\.256" /tmp/sass_loop.txt | head -1 | cut -d: -f1)
sed -n "$((L-8)),$((L+${2:-60}))p" /tmp/sass_loop.txt | cut -c1-100
cuobjdump -res-usage nrays_b200/csrc/libnrays_b200.so 2>/dev/null | grep -A1 "$pat" | grep -E "REG" | head -1
