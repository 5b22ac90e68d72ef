#!/usr/bin/env python
"""Top instructions of an `ncu --page source --csv` export by stall samples and by excessive
shared-memory wavefronts.  usage: ncu_source_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
print('total samples', tot, 'instructions', len(data))
print('--- by stall samples')
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:n]:
    stalls = {k[6:]: int(f(r, k)) for k in hdr if k.startswith('stall_') and f(r, k) > 0}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    print(f"{f(r,'# Samples')/tot*100:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {top}")
print('--- by excessive shared wavefronts')
for r in sorted(data, key=lambda r: -f(r, 'L1 Wavefronts Shared Excessive'))[:12]:
    if f(r, 'L1 Wavefronts Shared Excessive') <= 0: break
    print(f"{int(f(r,'L1 Wavefronts Shared Excessive')):10d} of {int(f(r,'L1 Wavefronts Shared')):10d}  {r[ix['Source']].strip()[:80]}")
