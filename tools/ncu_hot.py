#!/usr/bin/env python
"""Top stall sites from `ncu -i X.ncu-rep --page source --csv --kernel-id :::N` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[idx["# Samples"]].isdigit()]
tot = sum(int(r[idx["# Samples"]]) for r in data)
print("total samples", tot, "rows", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stalls}
print("stall mix:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    s = int(r[idx["# Samples"]])
    st = sorted(((int(r[idx[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f"{s:6d} {s / tot * 100:5.1f}% {r[idx['Source']][:100]:100s} {st}")
