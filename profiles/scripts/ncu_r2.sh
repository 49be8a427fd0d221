# Round 2: ncu launch list of the headline workload + one --set full capture of the classification kernel per workload and
# resident form (2 Mi fragments = 2 launches per step; the text legs come first in bench.py: 3 x 2 launches to skip).
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2_c4.csv \
    python bench.py --workloads c4 --steps 2 --warmup 1 --no-cpu-baseline --no-cli > gpurun_out/ncu_r2_l.log 2>&1
tail -4 gpurun_out/launches_r2_c4.csv
for w in c4 c3 c2; do
  for form in text packed; do
    skip=1; [ $form = packed ] && skip=7
    kern=analyze_reads_kernel
    # packed reads run the bulk kernel: its first launches are the packed legs'
    if [ $form = packed ]; then kern=analyze_bulk_kernel; skip=1; fi
    ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 1 -f \
        -o gpurun_out/prof_r2_${w}_${form} python bench.py --workloads $w --reads 2097152 --steps 1 --warmup 1 \
        --no-cpu-baseline --no-cli > gpurun_out/ncu_r2_${w}_${form}.log 2>&1
    python profiles/scripts/ncu_summary.py gpurun_out/prof_r2_${w}_${form}.ncu-rep > gpurun_out/prof_r2_${w}_${form}.txt 2>&1
    head -3 gpurun_out/prof_r2_${w}_${form}.txt
  done
done
python profiles/scripts/make_traffic.py gpurun_out
ls -la gpurun_out/*.ncu-rep
