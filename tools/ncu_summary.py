import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','launch__shared_mem_per_block_dynamic','launch__shared_mem_per_block_static','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_tensor.sum','launch__waves_per_multiprocessor']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:90])
    for i,h in enumerate(hdr):
        if h in keys: print(f'  {h} = {r[i]} {rows[1][i]}')
    st=[]
    for i,h in enumerate(hdr):
        if 'average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
            try: st.append((float(r[i]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
            except: pass
    print('  stalls:', ', '.join(f'{h}={v:.2f}' for v,h in sorted(st,reverse=True)[:8]))
