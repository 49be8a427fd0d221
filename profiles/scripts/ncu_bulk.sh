# Round 2, bulk kernel: one --set full capture of analyze_bulk_kernel per workload (packed form, SHK_BULK=1).
set -x
for w in ${WL:-c4}; do
  SHK_BULK=1 ncu --set full --clock-control none --import-source on -k regex:analyze_bulk_kernel -s 1 -c 1 -f \
      -o gpurun_out/prof_r2_${w}_bulk python bench.py --workloads $w --reads 2097152 --steps 1 --warmup 1 \
      --no-cpu-baseline --no-cli > gpurun_out/ncu_r2_${w}_bulk.log 2>&1
  python profiles/scripts/ncu_summary.py gpurun_out/prof_r2_${w}_bulk.ncu-rep > gpurun_out/prof_r2_${w}_bulk.txt 2>&1
  cat gpurun_out/prof_r2_${w}_bulk.txt
done
