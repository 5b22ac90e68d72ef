#!/usr/bin/env python
"""Print the interesting metrics of an `ncu --page raw --csv` export (one column per metric)."""
import csv, sys, re
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
  r'gpu__time_duration.sum|dram__bytes_(read|write)\.sum$|dram__throughput.avg.pct|sm__throughput.avg.pct|sm__warps_active.avg.pct|launch__registers|launch__occupancy_limit|achieved_occupancy|smsp__inst_executed.sum$|sm__inst_executed_pipe_(fma|fmaheavy|fmalite|alu|lsu|xu|uniform|fp64).*pct|smsp__average_warp.*stall|smsp__warp_issue_stalled.*ratio|l1tex__data_bank_conflicts|l1tex__throughput.avg.pct|lts__throughput.avg.pct|smsp__issue_active.avg.pct|sm__pipe.*cycles_active.avg.pct|shared_mem|smsp__average_warps_issue_stalled.*per_issue_active')
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
for r in data:
    print('==', r[hdr.index('Kernel Name')][:100], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for h, u, v in zip(hdr, units, r):
        if pat.search(h):
            print(f'  {h:100s} {v} {u}')
