#!/usr/bin/env python
"""Summarise ncu outputs for profiles/: (1) a launch list CSV (--metrics gpu__time_duration.sum) -> per-kernel totals and
shares; (2) a --set full report exported with `ncu -i X.ncu-rep --page raw --csv` -> key metrics per launch."""
import csv, sys, collections

def launches(path):
    rows = [r for r in csv.reader(open(path)) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
    tot = collections.OrderedDict()
    n = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) != len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        v = v / 1e3 if unit in ("ns", "nsecond") else v
        tot[name] = tot.get(name, 0.0) + v
        n[name] += 1
    s = sum(tot.values())
    print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"| `{k}` | {n[k]} | {v:.1f} | {v / s * 100:.1f} % |")
    print(f"| all | {sum(n.values())} | {s:.1f} | 100 % |")

def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"), ("launch__registers_per_thread", "regs"),
            ("launch__grid_size", "grid"), ("smsp__inst_executed.sum", "warp inst")]
    print("| id | kernel | " + " | ".join(w[1] for w in want) + " |\n|" + "---|" * (len(want) + 2))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("itg::", "")
        cells = []
        for m, _ in want:
            cells.append((r[idx[m]] + " " + units[idx[m]]).strip() if m in idx else "-")
        print(f"| {r[idx['ID']]} | `{name}` | " + " | ".join(cells) + " |")

if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else raw)(sys.argv[2])
