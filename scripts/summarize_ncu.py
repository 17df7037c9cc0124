import csv, sys, subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_xu.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__cycles_elapsed.max']
stalls=[h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    name=r[hdr.index('Kernel Name')]
    print('=====',name[:110])
    for w in want:
        if w in hdr:
            i=hdr.index(w); print('  %-75s %s %s'%(w,r[i],units[i]))
    st=sorted(((float(r[hdr.index(s)]),s) for s in stalls),reverse=True)[:7]
    for v,sn in st: print('  stall %-60s %.2f'%(sn.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v))
