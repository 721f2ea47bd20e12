#!/bin/bash
# usage: tools/sass_loop.sh <mangled kernel name> [context lines]  -- dumps SASS around the first MUFU.RCP (the substep loop)
SO=/root/repo/cartpolesimulation_b200/libcps_b200.so
K=$1; N=${2:-40}
cuobjdump -sass $SO | awk -v k="Function : $K" 'index($0,k){f=1;next} f&&/Function :/{f=0} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/\/\* 0x[0-9a-f]+ \*\/\s*$//' > /tmp/sass_k.txt
cuobjdump -res-usage $SO 2>/dev/null | grep -A1 "$K" | tail -1
L=$(grep -n "MUFU.RCP" /tmp/sass_k.txt | head -1 | cut -d: -f1)
sed -n "$((L-12)),$((L+N))p" /tmp/sass_k.txt | cut -c1-100
