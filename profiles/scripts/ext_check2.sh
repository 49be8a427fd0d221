set -x
for wl in c4 c3 c2; do bash profiles/scripts/ab.sh $wl EXTEND=off EXTEND=on; done > gpurun_out/ab_ext.log 2>&1; cat gpurun_out/ab_ext.log
# CLI stage timing on 8M C2 reads
python - <<PY
import sys; sys.path.insert(0,'.')
from shark_b200 import synth
names,bases,off=synth.make_reference(1000,seed=1)
synth.write_fasta('/tmp/ref.fa',names,bases,off)
seq,q,_=synth.make_reads(bases,1000,8000000,100,False,seed=2)
synth.write_fastq('/tmp/s_1.fq',None,seq,q,8000000,100,False)
PY
cd /tmp; for i in 1 2; do SHK_TIMING=1 $GRAFT_REPO_ROOT/shark_b200/shark-b200 -r ref.fa -1 s_1.fq -o g_1.fq > g.ssv 2> $GRAFT_REPO_ROOT/gpurun_out/cli_timing_$i.log; done; cat $GRAFT_REPO_ROOT/gpurun_out/cli_timing_2.log
