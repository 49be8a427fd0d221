# Validation of HEAD (v13) on one B200: all GPU tests, smoke, both bench arms (default run = what the driver does),
# C3/C4 lines, and the ncu launch list of the default bench command (split upload: the unpack kernel shows up in e2e).
set -x
nvidia-smi -L; nproc
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-400 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
for wl in c3 c4; do python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_$wl.json')); r=d['roofline']; e=d['e2e']; o=d['e2e_other']
print('$wl kernel_ms=%.4f value=%.1fM frac=%.3f e2e[%s]=%.1fM share=%.2f e2e[%s]=%.1fM build_ms=%.1f' % (r['kernel_ms_per_launch'], d['value']/1e6, r['frac'], e['upload'], e['value']/1e6, e['packed_share'], o['upload'], o['value']/1e6, d['index']['build_ms']))"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    k = r[ki].split('(')[0]; t[k][0] += 1; t[k][1] += v
tot = sum(v[1] for v in t.values())
for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]): print('%-60s n=%4d  %10.1f us  %5.1f %%' % (k[:60], v[0], v[1] / 1e3, 100 * v[1] / tot))
PY
