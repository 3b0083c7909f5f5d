"""Summarise an `ncu --page raw --csv` dump: python tools/ncu_summary.py raw.csv [row]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
vals = rows[2 + which]
d = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'lts__t_sectors.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'sm__cycles_elapsed.max',
        'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex_op_atom.sum', 'lts__t_sectors_srcunit_tex_op_red.sum']
for k in keys:
    if k in d:
        print('%-75s %s %s' % (k, d[k][0], d[k][1]))
print('--- warp stall reasons (warps per issue-active cycle, > 0.15)')
for h in hdr:
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
        try:
            v = float(d[h][0].replace(',', ''))
        except ValueError:
            continue
        if v > 0.15:
            print('  %-70s %.2f' % (re.sub(r'smsp__average_warps_issue_stalled_|_per_issue_active.ratio', '', h), v))
print('--- pct_of_peak (>= 20)')
for h in hdr:
    if h.endswith('pct_of_peak_sustained_elapsed') or h.endswith('pct_of_peak_sustained_active'):
        try:
            v = float(d[h][0].replace(',', ''))
        except ValueError:
            continue
        if v >= 20:
            print('  %-90s %.1f' % (h, v))
