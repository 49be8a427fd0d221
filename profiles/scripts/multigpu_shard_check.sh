# 2-GPU checks of the sharded index build: P2P OR-merge over NVLink between two processes (CUDA IPC, torchrun)
# and between two contexts of one process (peer access; pytest + the CLI with --gpus 2 --sharded-build).
set -x
nvidia-smi -L; nvidia-smi topo -m | head -6
( time timeout 300 python -m pytest tests/test_gpu_shard.py -x -q ) > gpurun_out/pytest_shard_2gpu.log 2>&1; tail -4 gpurun_out/pytest_shard_2gpu.log
for wl in c2 c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --workload $wl --reads 4000000 --index both --no-cpu-baseline > gpurun_out/bench_2gpu_shard_$wl.json 2> gpurun_out/bench_2gpu_shard_$wl.err
  echo rc=$?; grep -i "sharded\|index:\|error\|Traceback" gpurun_out/bench_2gpu_shard_$wl.err | tail -8
  python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu_shard_$wl.json')); print('$wl', {k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], d['index'])"
done
python - <<'PY'
import sys; sys.path.insert(0,'.')
from shark_b200 import synth
names,bases,off=synth.make_reference(200,seed=3)
synth.write_fasta('/tmp/ref.fa',names,bases,off)
seq,q,_=synth.make_reads(bases,200,600000,100,True,seed=5)
synth.write_fastq('/tmp/a_1.fq','/tmp/a_2.fq',seq,q,600000,100,True)
PY
cd /tmp
export SHK_TIMING=1
( time timeout 300 $GRAFT_REPO_ROOT/shark_b200/shark-b200 -r ref.fa -1 a_1.fq -2 a_2.fq -o g2_1.fq -p g2_2.fq --chunk-reads 100000 --gpus 2 > g2.ssv ) 2>&1 | grep -i "index\|real"
( time timeout 300 $GRAFT_REPO_ROOT/shark_b200/shark-b200 -r ref.fa -1 a_1.fq -2 a_2.fq -o h2_1.fq -p h2_2.fq --chunk-reads 100000 --gpus 2 --sharded-build > h2.ssv ) 2>&1 | grep -i "index\|real"
$GRAFT_REPO_ROOT/oracle/_ref/shark -r ref.fa -1 a_1.fq -2 a_2.fq -o r_1.fq -p r_2.fq -t 8 2>/dev/null | sort > r.sorted
wc -l r.sorted g2.ssv h2.ssv
sort g2.ssv | cmp - r.sorted && sort h2.ssv | cmp - r.sorted && cmp g2.ssv h2.ssv && cmp g2_1.fq h2_1.fq && cmp g2_2.fq h2_2.fq && echo CLI_SHARDED_BUILD_IDENTICAL
