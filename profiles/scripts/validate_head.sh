# Round-1 validation of the current tree on one B200: GPU tests, the default bench line, C3/C4 kernel times,
# the CLI on files against the reference binary, and the ncu launch list of the bench command.
set -x
nvidia-smi -L; nproc
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
bash profiles/scripts/ab.sh c3 "A=1" > gpurun_out/ab_c3.log 2>&1; cat gpurun_out/ab_c3.log
bash profiles/scripts/ab.sh c4 "A=1" > gpurun_out/ab_c4.log 2>&1; cat gpurun_out/ab_c4.log
bash profiles/scripts/cli_bench.sh 8000000 > gpurun_out/cli_bench.log 2>&1; tail -40 gpurun_out/cli_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/launches.csv
