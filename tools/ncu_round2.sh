set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.avg.per_second,launch__registers_per_thread,launch__grid_size,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.sum"
cap() { name=$1; shift; python tools/run_layer.py "$@" --iters 3 > gpurun_out/r2_ncu_${name}_plain.txt 2>&1; ncu --metrics $M --clock-control none -c 8 --csv --log-file gpurun_out/r2_ncu_${name}.csv python tools/run_layer.py "$@" --iters 3 > /dev/null 2>&1; tail -1 gpurun_out/r2_ncu_${name}_plain.txt; }
cap convt_dgrad_up4 --op convt_dgrad --cin 128 --cout 64 --hw 128 --batch 64
cap convt_dgrad_up3 --op convt_dgrad --cin 256 --cout 128 --hw 64 --batch 64
cap dgrad_bn_64 --op dgrad_bn --cin 64 --cout 64 --hw 256 --batch 64
cap dgrad_64 --op dgrad --cin 64 --cout 64 --hw 256 --batch 64
cap bilinear_fwd --op bilinear_fwd --cin 64 --hw 224 --batch 64
cap bilinear_bwd --op bilinear_bwd --cin 64 --hw 224 --batch 64
for op in head_fwd head_bwd bn_apply; do python tools/run_elem.py --op $op > gpurun_out/r2_elem_$op.txt 2>&1; tail -1 gpurun_out/r2_elem_$op.txt; done
