#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: hottest SASS instructions + stall mix (developer tool)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if 'Source' in r and '# Samples' in r)
body = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
i_src = hdr.index('Source'); i_s = hdr.index('# Samples'); i_ex = hdr.index('Instructions Executed')
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
tot = sum(int(r[i_s]) for r in body)
print('total samples', tot, 'instructions', len(body))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(body, key=lambda r: -int(r[i_s]))[:n]:
    st = {k: int(r[hdr.index(k)]) for k in stalls}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{int(r[i_s]):7d} {100 * int(r[i_s]) / tot:5.1f}% ex={r[i_ex]:>9s} {r[i_src].strip()[:72]:72s} {main}")
agg = {k: 0 for k in stalls}
for r in body:
    for k in stalls:
        agg[k] += int(r[hdr.index(k)])
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
