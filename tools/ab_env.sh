#!/bin/bash
# tools/ab_env.sh — dev helper: A/B the fused-chain bench across environment-variable settings of the same library,
# alternating runs on the same box. Usage: tools/ab_env.sh steps "VAR=1" "VAR=0 OTHER=3" ...
steps=${1:-30}; shift
cd "$(dirname "$0")/.."
run() { env $1 python bench.py --steps $steps --warmup 3 --no-e2e --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', round(d['value']/1e3,1), 'GS/s frac', round(d['roofline']['frac'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), d['clocks'])"; }
for rep in 1 2; do
  for cfg in "$@"; do run "$cfg"; done
done
