#!/usr/bin/env python3
"""Key metrics of an .ncu-rep (first kernel): python scripts/ncu_raw_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld_lookup_miss.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum', 'sm__cycles_elapsed.avg', 'launch__grid_size', 'launch__block_size']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u, v = rows[0], rows[1], rows[2]
d = {k: (v[i], u[i]) for i, k in enumerate(h)}
print(d.get('Kernel Name', ('?',))[0])
for k in KEYS:
    if k in d: print(f'  {k:70s} {d[k][0]:>16s} {d[k][1]}')
