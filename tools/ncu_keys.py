import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","smsp__thread_inst_executed_per_inst_executed.ratio","smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed","smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed","smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed","sm__cycles_elapsed.max","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__sass_inst_executed_op_local_ld.sum","smsp__sass_inst_executed_op_local_st.sum","dram__bytes_read.sum","dram__bytes_write.sum","launch__registers_per_thread"]
for i,h in enumerate(hdr):
    if h in keys or ("issue_stalled" in h and "per_issue_active" in h and float(vals[i] or 0) > 0.1):
        print(h, units[i], vals[i])
