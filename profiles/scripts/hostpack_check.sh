# Split upload (SHK_F_HOST_PACK): parity tests, then e2e plain vs split on C2 / C3 / C4 (one B200, 16 host cores).
set -x
nvidia-smi -L; nproc; lscpu | grep -i "model name"
( time python -m pytest tests/test_gpu_hostpack.py -x -q ) > gpurun_out/pytest_hostpack.log 2>&1; tail -6 gpurun_out/pytest_hostpack.log
show() { python -c "
import json,sys; d=json.loads(open('$1').read()); e=d['e2e']; o=d['e2e_other']
print('$1 value=%.1fM  e2e[%s]=%.1fM (%.1f MB h2d, %.2f ms)  e2e_other[%s]=%.1fM (%.1f MB h2d, %.2f ms)  pack=%s' % (d['value']/1e6, e['upload'], e['value']/1e6, e['h2d_bytes_per_step']/1e6, e['ms_per_step'], o['upload'], o['value']/1e6, o['h2d_bytes_per_step']/1e6, o['ms_per_step'], o['pack']))"; }
for wl in c2 c3 c4; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_split_$wl.json 2> gpurun_out/bench_split_$wl.err || tail -5 gpurun_out/bench_split_$wl.err
  show gpurun_out/bench_split_$wl.json
done
for share in 0.3 0.6 1.0; do
  timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --upload $share > gpurun_out/bench_split_c2_$share.json 2>/dev/null
  show gpurun_out/bench_split_c2_$share.json
done
SHK_PACK_THREADS=8 timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu-baseline --upload split > gpurun_out/bench_split_c2_t8.json 2>/dev/null; show gpurun_out/bench_split_c2_t8.json
