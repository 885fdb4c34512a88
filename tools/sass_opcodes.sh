#!/bin/bash
# SASS opcode histogram of the built library (what proves the Blackwell-native paths: UTMALDG/UTMASTG/UTMAREDG = tensor-map
# TMA loads / stores / reduce-adds, UBLKCP / UBLKRED = 1-D bulk copies / reduce-adds, SYNCS = mbarriers, USETMAXREG =
# register rebalancing between warpgroups).  No tensor-core opcodes (UTC*MMA) are expected: nothing on this path is a
# contraction.   usage: bash tools/sass_opcodes.sh > profiles/r2_sass_opcodes.txt
so=${1:-incompact3d_b200/libx3d_b200.so}
echo "# cuobjdump -sass $so | opcode histogram ($(git rev-parse --short HEAD 2>/dev/null), $(date -u +%F))"
cuobjdump -sass $so | grep -E "^\s+/\*[0-9a-f]{4}\*/" | awk '{print $2}' | sed 's/;$//' | sort | uniq -c | sort -rn
echo "# kernel entry points by template (instantiations):"
cuobjdump -sass $so | grep -E "Function :" | sed 's/.*Function : //' | c++filt | sed -E 's/^void //; s/<.*//; s/\(.*//; s/.*:://' | sort | uniq -c | sort -rn
