# Re-entry validation of HEAD on one B200: GPU tests, smoke, both bench arms, C3/C4 kernel times, and the
# ncu launch list + one full capture of the extension kernel on C4.
set -x
nvidia-smi -L; nproc
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
bash profiles/scripts/ab.sh c3 "A=1" > gpurun_out/ab_c3.log 2>&1; cat gpurun_out/ab_c3.log
bash profiles/scripts/ab.sh c4 "A=1" > gpurun_out/ab_c4.log 2>&1; cat gpurun_out/ab_c4.log
bash profiles/scripts/ncu_c4.sh > gpurun_out/ncu_c4.log 2>&1; tail -5 gpurun_out/ncu_c4.log
