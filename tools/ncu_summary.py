#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: one block per captured launch with the metrics the
profiles/ summaries quote."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = [hdr.index(w) for w in WANT if w in hdr]
for r in rows[2:]:
    print('---')
    for i in idx:
        print('%-85s %-12s %s' % (hdr[i], units[i], r[i][:110]))

# --json <workload> <out.json>: per kernel (first captured launch) the DRAM bytes and the time, for bench.py's roofline.traffic
if len(sys.argv) >= 5 and sys.argv[2] == '--json':
    import json
    import re
    col = {w: hdr.index(w) for w in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum') if w in hdr}
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    kernels = {}
    for r in rows[2:]:
        name = re.split(r'[<(]', r[col['Kernel Name']])[0].replace('void ', '').strip()
        if name in kernels:
            continue
        rd = float(r[col['dram__bytes_read.sum']].replace(',', '')) * scale.get(units[col['dram__bytes_read.sum']], 1.0)
        wr = float(r[col['dram__bytes_write.sum']].replace(',', '')) * scale.get(units[col['dram__bytes_write.sum']], 1.0)
        kernels[name] = {'dram_bytes_read': rd, 'dram_bytes_write': wr, 'gpu_time': r[col['gpu__time_duration.sum']] + ' ' + units[col['gpu__time_duration.sum']]}
    json.dump({'workload': sys.argv[3], 'source': 'ncu --set full --clock-control none, first captured launch of every kernel (tools/gpu_final.sh)', 'kernels': kernels},
              open(sys.argv[4], 'w'), indent=1)
