for MB in 4 6 8; do
  SHK_NVCC_FLAGS="-DSHK_FAST_MIN_BLOCKS=$MB" python shark_b200/build.py --force -v 2>&1 | grep -A2 "analyze_reads_kernelILb0ELi0" | tail -1
  SHK_FRONT_LOAD=0.7 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('minblocks $MB', d['value'], d['e2e']['value'], r['kernel_ms_per_launch'])"
done
