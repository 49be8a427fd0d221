#!/bin/bash
# A/B of fast-kernel variants on one box: prints kernel ms per 1 Mi-read launch, value and e2e.
# usage: ab.sh "<workload>" "<env assignments or BENCH_ARGS=...>" ...
wl=$1; shift
for cfg in "$@"; do
  extra=""
  case "$cfg" in EXTEND=*) extra="--extend ${cfg#EXTEND=}";; esac
  env $cfg timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline $extra 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl [$cfg] kernel_ms=%.4f value=%.1fM e2e=%.1fM frac=%.3f slow=%s extend=%s ext_frac=%.3f loads/probe=%.3f hit=%.3f build_ms=%.1f idx_GB=%.2f' % (r['kernel_ms_per_launch'], d['value']/1e6, d['e2e']['value']/1e6, r['frac'], d.get('slow_reads_per_step'), r['extend'], r['extended_fraction'], r['table_loads_per_probe'], r['hit_fraction'], d['index']['build_ms'], d['index']['device_bytes']/1e9))"
done
