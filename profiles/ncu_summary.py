import csv, sys, subprocess
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size',
 'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','sm__maximum_warps_per_active_cycle_pct']
for f in sys.argv[1:]:
    out = subprocess.run(['ncu','-i',f,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print('==', f, d['Kernel Name'][:60])
        for k in keys:
            if k in d: print(f'  {k:75s} {d[k]} {u[k]}')
        st = {h: float(d[h]) for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') or (h.startswith('smsp__average_warp') and 'per_issue_active' in h)}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:8]
        print('  stalls:', ', '.join(f"{k.split('issue_stalled_')[-1].split('_per_issue')[0]} {v:.1f}" for k, v in top))
