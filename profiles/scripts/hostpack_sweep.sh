# Split upload: auto-balanced share vs fixed shares on C2 / C3 / C4 (one B200).
show() { python -c "
import json,sys; d=json.loads(open('$1').read()); e=d['e2e']; o=d['e2e_other']
print('$2 upload=%s share=%.3f e2e=%.1fM (%.1f MB h2d, %.2f ms/step, pack %.1f Gbases/s) | plain=%.1fM | value=%.1fM | %s' % (e['upload'], e['packed_share'], e['value']/1e6, e['h2d_bytes_per_step']/1e6, e['ms_per_step'], e['pack_gbases_per_s'], o['value']/1e6, d['value']/1e6, o['pack']))"; }
run() { timeout 600 python bench.py --workload $1 --steps 8 --warmup 4 --no-cpu-baseline --upload $2 > gpurun_out/sweep_$1_$2.json 2>/dev/null; show gpurun_out/sweep_$1_$2.json $1; }
nproc
for u in split 0.5 0.7; do run c2 $u; done
for u in split 0.5 0.7 0.85; do run c3 $u; done
for u in split 0.5 0.65 0.8; do run c4 $u; done
