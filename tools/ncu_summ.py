import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r=list(csv.reader(out.splitlines()))
h=r[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warp_latency_issue_stalled_barrier.ratio']
for w in want:
    if w in h:
        i=h.index(w); print(w, r[1][i], [x[i] for x in r[2:]])
# stall reasons
for i,name in enumerate(h):
    if 'issue_stalled' in name and 'per_warp_active' in name and name.endswith('.pct'):
        print(name.replace('smsp__average_warps_issue_stalled_','').replace('_per_warp_active.pct',''), [x[i] for x in r[2:]])
