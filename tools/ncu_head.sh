# ncu of the fused head kernels (isolated launches at batch 64); each command runs plain first, then under ncu
set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,launch__grid_size,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active"
for op in head_fwd head_bwd; do
  python tools/run_elem.py --op $op --iters 3 > gpurun_out/r2_elem_${op}_v2.txt 2>&1 || exit 1
  tail -1 gpurun_out/r2_elem_${op}_v2.txt
  ncu --metrics $M --clock-control none -k regex:head -c 6 --csv --log-file gpurun_out/r2_ncu_${op}_mma.csv python tools/run_elem.py --op $op --iters 3 > /dev/null 2>&1
done
