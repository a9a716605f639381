#!/bin/bash
# tools/ncu_cfg5.sh — dev helper: one `ncu --set full` capture per recurrent-block kernel of config 5 (run under gpurun).
cd "$(dirname "$0")/.."
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --no-e2e --configs 5 --n5 67108864 --samples 16777216"
$B > gpurun_out/cfg5_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/cfg5_plain.log; exit 1; }
for k in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r02_$k $B > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
