#!/usr/bin/env python
"""Top stalled SASS instructions of an `ncu --page source --csv --print-source sass` dump (possibly several kernels)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None and len(r) > 10:
        cur['data'].append(r)
for b in blocks[:1]:
    hdr, data = b['hdr'], b['data']
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix['# Samples']]) for r in data)
    print(b['name'][:100], 'total samples', tot, 'instructions', len(data))
    top = sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n]
    for r in top:
        stalls = {k: int(r[ix[k]]) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k and r[ix[k]].isdigit() and int(r[ix[k]]) > 0}
        main = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        print(r[ix['# Samples']].rjust(7), r[ix['Source']][:64].ljust(64), r[ix['L1 Wavefronts Shared Excessive']].rjust(9), main)
