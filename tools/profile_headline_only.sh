cd "$(dirname "$0")/.." 2>/dev/null || cd /root/repo
O=gpurun_out; mkdir -p $O
FLOPS=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $FLOPS --clock-control none -k regex:"advance_kernel|setup_kernel|reduce_rows" -s 0 -c 12 --csv --log-file $O/r02_flops_headline.csv python bench.py --steps 1 --warmup 1 --no-extra --no-cpu > $O/r02_ncu_b2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:advance_kernel -s 2 -c 1 -f -o $O/r02_advance_headline python bench.py --steps 1 --warmup 1 --no-extra --no-cpu > $O/r02_ncu_b1.log 2>&1
ls -la $O/r02_advance_headline.ncu-rep $O/r02_flops_headline.csv
