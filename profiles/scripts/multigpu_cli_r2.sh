# Round 2, 2-GPU box: the two-process CLI with --gpus 2 (replicated and sharded index build), --wide-ids, gzip input, against the
# reference binary with -t 1 (identical bytes, identical order); then the GPU tests that spread contexts over both devices.
set -x
nvidia-smi -L
python - <<'PY'
import sys; sys.path.insert(0,'.')
from shark_b200 import synth
names,bases,off=synth.make_reference(200,seed=3)
synth.write_fasta('/tmp/ref.fa',names,bases,off)
seq,q,_=synth.make_reads(bases,200,600000,100,True,seed=5,varied_qual=True,want_qual=True)
synth.write_fastq('/tmp/a_1.fq','/tmp/a_2.fq',seq,q,600000,100,True)
PY
cd /tmp
gzip -k -1 a_1.fq a_2.fq
B=$GRAFT_REPO_ROOT/shark_b200/shark-b200
time $GRAFT_REPO_ROOT/oracle/_ref/shark -r ref.fa -1 a_1.fq -2 a_2.fq -o r_1.fq -p r_2.fq -q 20 -t 1 > r.ssv
time $B -r ref.fa -1 a_1.fq -2 a_2.fq -o g1_1.fq -p g1_2.fq -q 20 --chunk-reads 100000 > g1.ssv
time $B -r ref.fa -1 a_1.fq -2 a_2.fq -o g2_1.fq -p g2_2.fq -q 20 --chunk-reads 100000 --gpus 2 > g2.ssv
time $B -r ref.fa -1 a_1.fq -2 a_2.fq -o g3_1.fq -p g3_2.fq -q 20 --chunk-reads 70000 --gpus 2 --sharded-build > g3.ssv
time $B -r ref.fa -1 a_1.fq.gz -2 a_2.fq.gz -o g4_1.fq -p g4_2.fq -q 20 --gpus 2 > g4.ssv
time $B -r ref.fa -1 a_1.fq -2 a_2.fq -o g5_1.fq -p g5_2.fq -q 20 --wide-ids --gpus 2 | cat > g5.ssv
ok=1
for g in g1 g2 g3 g4 g5; do cmp $g.ssv r.ssv && cmp ${g}_1.fq r_1.fq && cmp ${g}_2.fq r_2.fq || ok=0; done
[ $ok = 1 ] && echo CLI_MULTI_GPU_IDENTICAL; wc -l r.ssv
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_shard.py tests/test_cli.py -m gpu -x -q 2>&1 | tail -3
