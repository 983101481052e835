#!/usr/bin/env python3
"""Per-instruction stall summary from `ncu -i rep --page source --csv` output (SASS view).
    ncu -i gpurun_out/prof.ncu-rep --page source --csv --kernel-name regex:conv1 --launch-count 1 > /tmp/src.csv
    python profiles/top_stalls.py /tmp/src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
idx = {name: i for i, name in enumerate(h)}
data = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break          # next launch's section
    if len(r) == len(h):
        data.append(r)
tot = sum(int(r[idx['# Samples']]) for r in data)
stall_cols = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
agg = {c: sum(int(r[idx[c]]) for r in data) for c in stall_cols}
print('total samples', tot, 'instructions', len(data))
print('by reason:', sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']]))[:n]:
    st = sorted(((int(r[idx[c]]), c[6:]) for c in stall_cols if int(r[idx[c]]) > 0), reverse=True)[:3]
    print(r[idx['# Samples']].rjust(6), r[idx['Instructions Executed']].rjust(8), r[idx['Address']][-5:], r[idx['Source']].strip()[:72].ljust(72), st)
