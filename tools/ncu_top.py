#!/usr/bin/env python
"""Top stalled SASS instructions from `ncu --page source --csv` output."""
import csv, sys
path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = [r for r in csv.reader(open(path))]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
tot = sum(int(r[idx['# Samples']]) for r in data)
print('total samples', tot, 'instructions', len(data))
seq = {id(r): i for i, r in enumerate(data)}
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:n]:
    st = {k[6:]: int(r[idx[k]]) for k in hdr if k.startswith('stall_') and '(Not' not in k and r[idx[k]] not in ('', '0')}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(str(seq[id(r)]).rjust(5), r[idx['# Samples']].rjust(6), r[idx['Instructions Executed']].rjust(10), r[idx['Source']].strip()[:64].ljust(64), st)
