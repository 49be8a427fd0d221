# 2-GPU checks: bench.py under torchrun (NCCL broadcast of the index) and the CLI with --gpus 2
set -x
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -5 gpurun_out/bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['e2e'], d['index'])"
# CLI: 2 GPUs vs 1 GPU vs reference on a generated sample
python - <<'PY'
import sys; sys.path.insert(0,'.')
from shark_b200 import synth
names,bases,off=synth.make_reference(200,seed=3)
synth.write_fasta('/tmp/ref.fa',names,bases,off)
seq,q,_=synth.make_reads(bases,200,600000,100,True,seed=5)
synth.write_fastq('/tmp/a_1.fq','/tmp/a_2.fq',seq,q,600000,100,True)
PY
cd /tmp
time $GRAFT_REPO_ROOT/shark_b200/shark-b200 -r ref.fa -1 a_1.fq -2 a_2.fq -o g1_1.fq -p g1_2.fq --chunk-reads 100000 > g1.ssv
time $GRAFT_REPO_ROOT/shark_b200/shark-b200 -r ref.fa -1 a_1.fq -2 a_2.fq -o g2_1.fq -p g2_2.fq --chunk-reads 100000 --gpus 2 > g2.ssv
time $GRAFT_REPO_ROOT/oracle/_ref/shark -r ref.fa -1 a_1.fq -2 a_2.fq -o r_1.fq -p r_2.fq -t 1 > r.ssv
cmp g1.ssv r.ssv && cmp g2.ssv r.ssv && cmp g1_1.fq r_1.fq && cmp g2_1.fq r_1.fq && cmp g2_2.fq r_2.fq && echo CLI_MULTI_GPU_IDENTICAL; wc -l r.ssv
