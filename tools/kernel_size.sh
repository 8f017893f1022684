#!/bin/bash
# SASS instruction count, registers and spills of the timed Whitted kernel in a library: tools/kernel_size.sh <lib.so>
K='_ZN3rtb14whitted_kernelILb0ELi3ELi3ELb0ELb0ELb1EEEvNS_6WFrameENS_5ShardEjPKjS4_jPjS5_PyNS_5PtBvhEPKhj'
n=$(cuobjdump -sass -fun "$K" "$1" 2>/dev/null | grep -cE "^\s+/\*[0-9a-f]{4,5}\*/")
l=$(cuobjdump -sass -fun "$K" "$1" 2>/dev/null | grep -cE "STL|LDL")
r=$(cuobjdump -res-usage "$1" 2>/dev/null | grep -A1 "$K" | grep -oE "REG:[0-9]+ STACK:[0-9]+")
echo "$(basename $1): $n instructions ($((n*16/1024)) KB), $l local-memory instructions, $r"
