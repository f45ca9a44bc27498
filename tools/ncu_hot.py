#!/usr/bin/env python
"""Top SASS lines of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass` output: by instructions
executed and by stall samples.  Profiling aid (the .ncu-rep with embedded source is too large to bring back).
    python tools/ncu_hot.py sass.csv [n]"""
import csv, sys
path, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, errors='replace')))
hdr_i = next(i for i, r in enumerate(rows) if any('Instructions Executed' in c for c in r))
hdr = rows[hdr_i]
print('columns:', hdr)
col = lambda name: next(i for i, c in enumerate(hdr) if name in c)
ci, cs, src = col('Instructions Executed'), col('Sampling'), col('Source')
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
num = lambda x: float(x.replace(',', '') or 0) if x.replace(',', '').replace('.', '').isdigit() else 0.0
tot_i, tot_s = sum(num(r[ci]) for r in body), sum(num(r[cs]) for r in body)
print(f'total warp-instructions {tot_i:.0f}, stall samples {tot_s:.0f}, sass lines {len(body)}')
for key, c in (('instructions', ci), ('samples', cs)):
    print(f'--- top {n} by {key}')
    for r in sorted(body, key=lambda r: -num(r[c]))[:n]:
        print(f'{num(r[ci]):9.0f} {num(r[cs]):7.0f}  {r[src][:110]}')
