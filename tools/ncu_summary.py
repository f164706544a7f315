import csv,sys,subprocess,io
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=rows[0]; units=rows[1]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_write.sum','lts__t_sectors_srcunit_tex_op_atom.sum','lts__t_sectors_srcunit_tex_op_red.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','smsp__inst_executed_op_global_atom.sum','smsp__inst_executed_op_global_red.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print('==',d['Kernel Name'][:70], d.get('Grid Size'), d.get('Block Size'))
    for k in keys:
        if k in d: print('  ',k,d[k],units[hdr.index(k)])
    st=[]
    for k in hdr:
        if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio'):
            try: v=float(d[k].replace(',',''))
            except: continue
            st.append((v,k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
    print('   stalls/issue:', ' '.join(f'{n}={v:.2f}' for v,n in sorted(st,reverse=True)[:8]))
