#!/usr/bin/env python
"""ncu --metrics gpu__time_duration.sum --csv log -> per-kernel launch list (json on stdout).
usage: launch_list.py <launches.csv> "<command that was profiled>" """
import csv, json, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iK, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iV].replace(",", ""))
    us = v / 1000.0 if r[iU] in ("ns", "nsecond") else (v if r[iU] in ("us", "usecond") else v * 1000.0)
    name = r[iK].split("(")[0].replace("void ", "")
    d = acc.setdefault(name, {"launches": 0, "total_us": 0.0})
    d["launches"] += 1
    d["total_us"] += us
tot = sum(d["total_us"] for k, d in acc.items() if "pnb::" in k or k.startswith("k_"))
for k, d in acc.items():
    d["total_us"] = round(d["total_us"], 1)
    d["avg_us"] = round(d["total_us"] / d["launches"], 1)
    if "pnb::" in k or k.startswith("k_"):
        d["share_of_pnb_time"] = round(d["total_us"] / tot, 4)
print(json.dumps({"command": sys.argv[2], "note": "cold-cache, serialised per-launch times; compare SHARES. "
                  "Includes the untimed setup sweeps (CountCl), the first CSR build and the e2e legs of bench.py.",
                  "kernels": acc}, indent=1))
