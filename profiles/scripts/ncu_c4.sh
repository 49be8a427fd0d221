# ncu: launch list + one full capture of the extension kernel on C4 (2 Mi pairs = 2 launches per step)
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --reads 2097152 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c4_l.log 2>&1
tail -8 gpurun_out/launches_c4.csv
ncu --set full --clock-control none --import-source on -k regex:analyze_reads_kernel -s 3 -c 1 -o gpurun_out/prof_c4_ext -f python bench.py --workload c4 --reads 2097152 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c4_f.log 2>&1
tail -3 gpurun_out/ncu_c4_f.log; ls -la gpurun_out/
