set -x
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q ) > gpurun_out/pytest_ext.log 2>&1; tail -5 gpurun_out/pytest_ext.log
for wl in c4 c3; do bash profiles/scripts/ab.sh $wl EXTEND=on SHK_LIB=$PWD/shark_b200/libshark_b200_mb6.so; done > gpurun_out/ab_ext2.log 2>&1; cat gpurun_out/ab_ext2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --reads 2097152 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c4_l.log 2>&1
grep -E "analyze_(mid|slow|reads)" gpurun_out/launches_c4.csv | tail -6
