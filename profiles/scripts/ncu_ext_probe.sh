# Round 1, v11: default bench line (with the stand-alone probe object), then one full ncu capture each of
#   - the extension kernel on C4 (analyze_reads_kernel<.., EXT>),
#   - the stand-alone probe kernel on C4 (probe_bench_kernel: hash -> filter word -> sector rank -> entry),
# plus the C4 launch list.  ncu numbers are cold-cache and serialised; bench values come from the plain run.
set -x
python bench.py > gpurun_out/bench_c2_v11.json 2> gpurun_out/bench_c2_v11.err; tail -2 gpurun_out/bench_c2_v11.err; cut -c1-400 gpurun_out/bench_c2_v11.json
python bench.py --workload c4 --reads 4194304 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_v11.json 2> gpurun_out/bench_c4_v11.err; tail -2 gpurun_out/bench_c4_v11.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_v11.json')); print(d['value'], d['e2e']['value'], d['roofline'], d['probe'])"
ncu --set full --clock-control none --import-source on -k regex:analyze_reads_kernel -s 3 -c 1 -o gpurun_out/prof_c4_ext_v11 -f python bench.py --workload c4 --reads 2097152 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c4_ext.log 2>&1
tail -2 gpurun_out/ncu_c4_ext.log
ncu --set full --clock-control none --import-source on -k regex:probe_bench_kernel -s 2 -c 1 -o gpurun_out/prof_c4_probe_v11 -f python bench.py --workload c4 --reads 1048576 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c4_probe.log 2>&1
tail -2 gpurun_out/ncu_c4_probe.log
ls -la gpurun_out/*.ncu-rep
