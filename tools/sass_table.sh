#!/bin/bash
# tools/sass_table.sh — SASS opcode histogram of one kernel of libqdsp_b200.so (offline, cuobjdump).
# Usage: tools/sass_table.sh <object file under qdsp_b200/csrc/build> <mangled-name substring> > profiles/rNN_sass_<kernel>.txt
obj=$1; pat=$2
fn=$(cuobjdump -sass "$obj" | grep "Function :" | grep "$pat" | head -1 | awk '{print $3}')
echo "kernel: $fn ($(echo $fn | c++filt))"
echo "object: $obj (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo)"
cuobjdump -sass -fun "$fn" "$obj" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\*[0-9a-fx ]*\*/##g' | awk '{ if ($1 ~ /^@/) print $2; else print $1}' | sed 's/;$//' | sort | uniq -c | sort -rn
