import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__shared_mem_per_block_dynamic','sm__warps_active.avg.pct_of_peak_sustained_active',
 'sm__cycles_elapsed.max','smsp__inst_executed.sum','sm__inst_executed.sum.per_cycle_elapsed','sm__issue_active.avg.pct_of_peak_sustained_elapsed',
 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'lts__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__average_warp_latency_per_inst_issued.ratio']
stall = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and '_per_issue_active' in h]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in keys:
        for h in hdr:
            if h == k: print(f'{k} [{units[hdr.index(h)]}] = {d[h]}')
    st = sorted(((float(d[h] or 0), h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for h in stall), reverse=True)
    print('stalls:', ', '.join(f'{n}={v:.2f}' for v,n in st[:9]))
