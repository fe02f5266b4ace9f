#!/usr/bin/env python
"""Prints dram bytes / duration per profiled launch from an `ncu --csv --metrics ...` log."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
out = {}
for r in rows[hdr + 1:]:
    d = dict(zip(h, r))
    out.setdefault(int(d["ID"]), {"kernel": d["Kernel Name"][:40]})[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
for k in sorted(out):
    v = out[k]
    print(k, v["kernel"], " ".join(f"{m}={x:.4g}" for m, x in v.items() if m != "kernel"))
