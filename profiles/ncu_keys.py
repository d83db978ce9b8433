#!/usr/bin/env python
"""Print the headline metrics of `ncu --page raw --csv` exports side by side: python profiles/ncu_keys.py a_raw.csv [b_raw.csv ...]"""
import csv
import sys

KEYS = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__registers_per_thread', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_active.avg', 'sm__cycles_active.max', 'sm__cycles_active.min']
cols = []
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        d['_units'] = dict(zip(hdr, units))
        cols.append(d)
keys = KEYS + sorted({h for c in cols for h in c if 'issue_stalled' in h and 'per_issue_active' in h})
for k in keys:
    vals = [c.get(k, '-') for c in cols]
    if all(v == '-' for v in vals):
        continue
    name = k.replace('smsp__average_warps_issue_stalled_', 'stall:').replace('_per_issue_active.ratio', '')
    print('%-62s %-8s %s' % (name[:62], cols[0]['_units'].get(k, '')[:8], '  '.join(v[:34] for v in vals)))
