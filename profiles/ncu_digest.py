#!/usr/bin/env python
"""Digest an `ncu --page raw --csv` dump (+ optional `--page source --csv`) into the few numbers we track."""
import csv, sys
raw = sys.argv[1]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
for h, u, v in zip(hdr, units, vals):
    if h in keep or ('issue_stalled' in h and 'per_issue_active' in h and float(v or 0) > 0.5):
        print(f"{h} [{u}] = {v}")
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    hdr = rows[1]
    ia, isamp, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    data = [(int(r[isamp]), int(r[iex]), idx, r[ia].strip()) for idx, r in enumerate(rows[2:]) if len(r) > iex and r[isamp].isdigit()]
    tot = sum(d[0] for d in data)
    print(f"-- top SASS by stall samples (total {tot}, {sum(d[1] for d in data)} warp instructions)")
    for d in sorted(data, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 20]:
        print(f"{d[0]:7d} {100*d[0]/tot:5.1f}%  ex={d[1]:9d}  #{d[2]:4d} {d[3][:100]}")
