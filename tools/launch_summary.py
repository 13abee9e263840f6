#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
usage: python tools/launch_summary.py gpurun_out/launches.csv "command line" > profiles/rN_launches.txt"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)
        a = agg.setdefault(name, dict(n=0, us=0.0, grid=r["Grid Size"], block=r["Block Size"]))
        a["n"] += 1
        a["us"] += us
    total = sum(a["us"] for a in agg.values())
    if len(sys.argv) > 2:
        print(sys.argv[2])
    print("(per-launch times are cold-cache and serialised: compare SHARES)")
    print(f"{'kernel':44s} {'launches':>8s} {'avg us':>9s} {'share':>6s}  grid / block")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{name[:44]:44s} {a['n']:8d} {a['us'] / a['n']:9.1f} {a['us'] / total:6.3f}  {a['grid']} / {a['block']}")


if __name__ == "__main__":
    main()
