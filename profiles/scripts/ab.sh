#!/bin/bash
# A/B of fast-kernel variants on one box: prints kernel ms per 1 Mi-read launch, value and e2e.
# usage: ab.sh "<workload>" "<env assignments>" ...
wl=$1; shift
for cfg in "$@"; do
  env $cfg python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl [$cfg] kernel_ms=%.4f value=%.1fM e2e=%.1fM frac=%.3f slow=%s' % (r['kernel_ms_per_launch'], d['value']/1e6, d['e2e']['value']/1e6, r['frac'], d.get('slow_reads_per_step')))"
done
