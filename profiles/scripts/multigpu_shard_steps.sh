# 2 GPUs, one process each: where the wall time of the sharded index build goes (per-step host clock), and the
# 2-GPU bench line with the split upload (pack threads shared between the ranks).
set -x
nproc
for wl in c2 c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload $wl --reads 4000000 --index both --no-cpu-baseline > gpurun_out/bench_2gpu_steps_$wl.json 2> gpurun_out/bench_2gpu_steps_$wl.err
  echo rc=$?; grep -i "sharded\|index:\|error\|Traceback" gpurun_out/bench_2gpu_steps_$wl.err | tail -6
  python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu_steps_$wl.json')); print('$wl', 'value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, d['e2e']['upload'], d['e2e']['packed_share'], 'plain', d['e2e_other']['value']/1e6, d['e2e_other']['pack']); print(d['index'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_v13.json 2> gpurun_out/bench_2gpu_v13.err; tail -2 gpurun_out/bench_2gpu_v13.err; cat gpurun_out/bench_2gpu_v13.json
