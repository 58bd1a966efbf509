"""Print the headline metrics of an ncu report (raw page). usage: ncu_keys.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_global_red.sum']
for h, u, v in zip(hdr, units, vals):
    if h in want or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
        print('%-80s %-12s %s' % (h, u, v))
