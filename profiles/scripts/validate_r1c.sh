# Round-1 re-entry (session 8): sharded index build + P2P OR-merge, index save/load, segment-timed build_ms.
# One B200: the new GPU tests (contexts share device 0), memcheck of the sharded path, CLI save/load against
# the reference binary, C2/C3/C4 build times.
set -x
nvidia-smi -L; nproc
( time python -m pytest tests/test_gpu_shard.py -x -q ) > gpurun_out/pytest_shard.log 2>&1; tail -15 gpurun_out/pytest_shard.log
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_shard.py -x -q -k "degenerate or step_by_step or limits" ) > gpurun_out/memcheck_shard.log 2>&1; echo memcheck rc=$?; tail -6 gpurun_out/memcheck_shard.log
# CLI: --save-index / --load-index give the reference's bytes
python - <<'PY'
import sys; sys.path.insert(0,'.')
from shark_b200 import synth
names,bases,off=synth.make_reference(200,seed=3)
synth.write_fasta('/tmp/ref.fa',names,bases,off)
seq,q,_=synth.make_reads(bases,200,300000,100,True,seed=5)
synth.write_fastq('/tmp/a_1.fq','/tmp/a_2.fq',seq,q,300000,100,True)
PY
( cd /tmp
export SHK_TIMING=1
$GRAFT_REPO_ROOT/shark_b200/shark-b200 -r ref.fa -1 a_1.fq -2 a_2.fq -o s_1.fq -p s_2.fq --save-index idx.shk > s.ssv 2> s.err; grep -i "index" s.err
$GRAFT_REPO_ROOT/shark_b200/shark-b200 -r ref.fa -1 a_1.fq -2 a_2.fq -o l_1.fq -p l_2.fq --load-index idx.shk > l.ssv 2> l.err; grep -i "index" l.err
$GRAFT_REPO_ROOT/oracle/_ref/shark -r ref.fa -1 a_1.fq -2 a_2.fq -o r_1.fq -p r_2.fq -t 1 > r.ssv 2>/dev/null
ls -l idx.shk; cmp s.ssv r.ssv && cmp l.ssv r.ssv && cmp l_1.fq r_1.fq && cmp l_2.fq r_2.fq && cmp s_2.fq r_2.fq && echo CLI_SAVE_LOAD_IDENTICAL; wc -l r.ssv ) 2>&1 | tee gpurun_out/cli_save_load.log
for wl in c2 c3 c4; do bash profiles/scripts/ab.sh $wl "A=1" "A=2"; done > gpurun_out/ab_build.log 2>&1; cat gpurun_out/ab_build.log
