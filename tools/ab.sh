#!/bin/bash
# tools/ab.sh — dev helper: A/B the fused-chain bench between the in-tree library and alternative builds
# (qdsp_b200/csrc/build/ab/*.so), alternating runs on the same box. Usage: tools/ab.sh [steps]
steps=${1:-30}
cd "$(dirname "$0")/.."
cp qdsp_b200/libqdsp_b200.so /tmp/cur.so
run() { python bench.py --steps $steps --warmup 3 --no-e2e --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', round(d['value']/1e3,1), 'GS/s frac', round(d['roofline']['frac'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), d['clocks'])"; }
for rep in 1 2; do
  for alt in qdsp_b200/csrc/build/ab/*.so; do
    cp "$alt" qdsp_b200/libqdsp_b200.so; run "$(basename $alt)"
  done
  cp /tmp/cur.so qdsp_b200/libqdsp_b200.so; run current
done
