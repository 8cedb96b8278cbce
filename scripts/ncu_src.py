#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: top SASS/source lines by stall samples.
usage: python scripts/ncu_src.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# find header row
h = next(i for i, r in enumerate(rows) if r and r[0] in ("Address", "#"))
hdr = rows[h]
idx = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
def num(x):
    try: return float(x)
    except: return 0.0
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(num(r[idx["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {s: sum(num(r[idx[s]]) for r in data) for s in stalls}
print("stall totals:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0})
data.sort(key=lambda r: -num(r[idx["# Samples"]]))
for r in data[:top]:
    st = {s[6:]: int(num(r[idx[s]])) for s in stalls if num(r[idx[s]]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f'{int(num(r[idx["# Samples"]])):6d} {100*num(r[idx["# Samples"]])/max(tot,1):5.1f}%  ex={int(num(r[idx["Instructions Executed"]])):9d}  {r[idx["Source"]][:90]:90s} {st}')
