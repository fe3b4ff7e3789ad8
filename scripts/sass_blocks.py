#!/usr/bin/env python
"""dev helper: bucket the SASS of an ncu report (--page source --csv) into basic blocks by execution count."""
import csv, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
I = lambda r, k: int(r[ix[k]] or 0)
tot = sum(I(r, 'Instructions Executed') for r in data); samp = sum(I(r, '# Samples') for r in data)
print('total inst', tot, 'samples', samp)
names = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
blk = []; cur = None
for i, r in enumerate(data):
    n = I(r, 'Instructions Executed'); s = I(r, '# Samples')
    brk = any(t in r[1] for t in ('RET', 'EXIT'))
    if cur and abs(n - cur['n']) <= 0.02 * max(n, cur['n'], 1):
        cur['cnt'] += 1; cur['inst'] += n; cur['s'] += s; cur['end'] = i
    else:
        cur = {'n': n, 'cnt': 1, 'inst': n, 's': s, 'start': i, 'end': i}; blk.append(cur)
    if brk: print('  -- fn boundary at', i, r[1].strip()[:40]); cur = None
for b in blk:
    if b['inst'] > thr * tot or b['s'] > thr * samp:
        d = {k: sum(I(r, k) for r in data[b['start']:b['end'] + 1]) for k in names}
        top = sorted(d.items(), key=lambda x: -x[1])[:3]
        print(f"{b['start']:5d}-{b['end']:5d} n={b['n']:>10d} len={b['cnt']:4d} inst%={100*b['inst']/tot:5.1f} samp%={100*b['s']/samp:5.1f} {top} | {data[b['start']][1].strip()[:40]}")
if len(sys.argv) > 4:
    for i in range(int(sys.argv[3]), int(sys.argv[4]) + 1):
        print(i, data[i][1].strip()[:80], data[i][ix['Instructions Executed']], data[i][ix['# Samples']])
