import csv, sys, subprocess
for f in sys.argv[1:]:
    out = subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print('==', f, d.get('Kernel Name'))
        keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
                'smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
                'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
                'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
                'smsp__sass_inst_executed_op_local_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
                'smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum', 'sm__cycles_active.avg']
        for k in keys:
            if k in d: print(f'  {k} = {d[k]} {units[hdr.index(k)]}')
        st = {k.replace('smsp__pcsamp_warps_issue_stalled_',''): int(v) for k,v in d.items() if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued') and v.isdigit()}
        tot = sum(st.values())
        print('  stalls:', ', '.join(f'{k}={v*100//tot}%' for k,v in sorted(st.items(), key=lambda x:-x[1])[:8]))
