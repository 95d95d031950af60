import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.sum','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum','smsp__cycles_active.avg']
extra=sys.argv[2:] 
for r in rows[2:]:
  print('---')
  for k in keys+extra:
    for i,h in enumerate(hdr):
        if h==k: print(k,'=',r[i], rows[1][i])
