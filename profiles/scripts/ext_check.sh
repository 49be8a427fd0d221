# anchor-and-extend: parity (forced on small inputs, automatic at C3/C4 size) and A/B kernel times
set -x
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q ) > gpurun_out/pytest_ext.log 2>&1; tail -15 gpurun_out/pytest_ext.log
for wl in c4 c3 c2; do bash profiles/scripts/ab.sh $wl EXTEND=off EXTEND=on; done > gpurun_out/ab_ext.log 2>&1; cat gpurun_out/ab_ext.log
