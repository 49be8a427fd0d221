#!/bin/bash
# CLI end to end on files: shark-b200 vs the reference binary on the same FASTQ (C2-shaped, N reads).
N=${1:-8000000}
python - <<PY
import sys; sys.path.insert(0,'.')
from shark_b200 import synth
names,bases,off=synth.make_reference(1000,seed=1)
synth.write_fasta('/tmp/ref.fa',names,bases,off)
seq,q,_=synth.make_reads(bases,1000,$N,100,False,seed=2)
synth.write_fastq('/tmp/s_1.fq',None,seq,q,$N,100,False)
PY
ls -la /tmp/s_1.fq
cd /tmp
for i in 1 2 3; do
echo shark-b200; time $GRAFT_REPO_ROOT/shark_b200/shark-b200 -r ref.fa -1 s_1.fq -o g_1.fq $CLI_FLAGS > g.ssv
done
echo reference -t $(nproc); time $GRAFT_REPO_ROOT/oracle/_ref/shark -r ref.fa -1 s_1.fq -o r_1.fq -t $(nproc) > r.ssv
sort g.ssv | md5sum; sort r.ssv | md5sum; wc -l g.ssv r.ssv
# FASTQ records keyed by name
paste - - - - < g_1.fq | sort | md5sum; paste - - - - < r_1.fq | sort | md5sum
