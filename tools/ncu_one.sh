#!/bin/bash
# tools/ncu_one.sh — dev helper (under gpurun): one `ncu --set full` capture of the first launch of kernel <regex> in <command...>
# Usage: tools/ncu_one.sh <name> <kernel regex> <command...>   -> gpurun_out/<name>.ncu-rep
cd "$(dirname "$0")/.."
name=$1; k=$2; shift; shift
ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/$name "$@" > gpurun_out/ncu_$name.log 2>&1
echo "$name rc=$?"
