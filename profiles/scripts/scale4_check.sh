# 4-GPU check of the driver's scaling launch: bench.py under torchrun at N=4 (C2 per rank, index built on
# rank 0 and broadcast with NCCL), and a C4 run (10 GB index incl. the extension structures) at N=4.
set -x
nvidia-smi -L | head -8; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err
tail -4 gpurun_out/bench_4gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_4gpu.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['e2e'], d['index'], d['clocks'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --workload c4 --reads 4194304 --steps 3 --warmup 3 > gpurun_out/bench_4gpu_c4.json 2> gpurun_out/bench_4gpu_c4.err
tail -4 gpurun_out/bench_4gpu_c4.err
python -c "
import json; d=json.load(open('gpurun_out/bench_4gpu_c4.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['e2e'], d['index'])"
