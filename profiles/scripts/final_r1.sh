# Last check of HEAD (v13): split-upload + sharded-build tests, the default bench line, and one ncu --set full capture of
# the unpack kernel (the only kernel the split upload adds to a step).
set -x
( python -m pytest tests/test_gpu_hostpack.py tests/test_gpu_shard.py -q 2>&1 | tail -2 )
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -1 gpurun_out/bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); e=d['e2e']; print('value %.1fM e2e %.1fM share %.2f pack %.1f plain %.1fM parity %s frac %.3f' % (d['value']/1e6, e['value']/1e6, e['packed_share'], e['pack_gbases_per_s'], d['e2e_other']['value']/1e6, d['cpu_baseline']['parity_on_sample'], d['roofline']['frac']))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:unpack_reads_kernel -s 5 -c 1 -o gpurun_out/prof_unpack -f python bench.py --reads 3145728 --steps 1 --warmup 1 --no-cpu-baseline --upload 1.0 > gpurun_out/ncu_unpack.log 2>&1; tail -2 gpurun_out/ncu_unpack.log; ls -la gpurun_out/prof_unpack.ncu-rep
