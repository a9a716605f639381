#!/bin/bash
# tools/ab_cfg.sh — dev helper: run one of bench.py's secondary configs under several environment settings.
# Usage: tools/ab_cfg.sh <config key, e.g. cfg4> "<extra bench args>" "ENV=1" "ENV=2 OTHER=3" ...
key=$1; shift; extra=$1; shift
cd "$(dirname "$0")/.."
cfgnum=$(echo $key | sed 's/cfg\([0-9]\).*/\1/')
for e in "$@"; do
  env $e python bench.py --steps 3 --warmup 3 --no-cpu --no-parity --no-e2e --configs $cfgnum $extra 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        for k,c in d['configs'].items():
            if k.startswith('$key'):
                r=c['roofline']; print('$e', k, round(c['value'],1), 'frac', round(r.get('frac_nominal', r.get('frac', 0)),4), 'parity', c['parity']['ok'], c['clocks']['sm_mhz'], c['clocks']['reasons'])
"
done
