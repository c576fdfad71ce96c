#!/usr/bin/env python
"""Per-kernel times of the LAST sketch batch in an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); iid = hdr.index('ID')
L = [(int(r[iid]), r[ik], float(r[iv].replace(',', '')) / 1000) for r in rows[1:] if len(r) > iv]
idx = [i for i, (a, k, t) in enumerate(L) if 'scan_' in k]
st = idx[-1] if idx else 0
tot = 0
for a, k, t in L[st:st + 80]:
    print(f"{t:9.1f} us  {re.sub(r'\(.*', '', k)[:90]}")
    tot += t
    if 'hashjoin' in k:
        break
print('total', round(tot, 1), 'us over', len(L), 'launches in the file')
