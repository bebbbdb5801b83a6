#!/bin/bash
# usage: scripts/sass_loop.sh lib.so  — prints the first inner node loop of trace_kernel<false,false>
F='_ZN3nrb12trace_kernelILb0ELb0EEEvNS_9SceneViewENS_11FrameParamsENS_8RayQueueEP6float4PNS_12WaveCountersEjjNS_11ShadowQueueES5_S7_iii'
cuobjdump -sass -fun "$F" "$1" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##' | cut -c1-100
