# Round 2: bulk kernel (K6b) variants on one box - parity tests once, then kernel ms / value per variant library.
# usage: bulk_sweep.sh "<workloads>" <variant names...>   (libshark_b200_<name>.so built by build.py --variant)
set -x
wls=$1; shift
timeout 900 python -m pytest tests/test_gpu_bulk.py -x -q > gpurun_out/bulk_tests.log 2>&1; tail -3 gpurun_out/bulk_tests.log
for wl in $wls; do
SHK_BULK=0 timeout 600 python bench.py --workloads $wl --steps 5 --warmup 3 --no-cpu-baseline --no-cli 2>/dev/null > gpurun_out/bulk_${wl}_base.json
for v in "$@"; do
  SHK_BULK=1 SHK_LIB=$PWD/shark_b200/libshark_b200_$v.so timeout 600 python bench.py --workloads $wl --steps 5 --warmup 3 --no-cpu-baseline --no-cli 2>/dev/null > gpurun_out/bulk_${wl}_$v.json
done
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/bulk_${wl}_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print('%-40s kernel_ms=%.4f value=%.1fM e2e=%.1fM e2e_packed=%s ext=%.3f loads/probe=%.3f' % (f, r['kernel_ms_per_launch'], d['value']/1e6, d['e2e']['value']/1e6, d.get('e2e_packed',{}).get('value'), r['extended_fraction'], r['table_loads_per_probe']))
    except Exception as e: print(f, 'failed', e)
P
done
