"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-launch and per-kernel-class times."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
tot, cls = 0.0, OrderedDict()
for i, r in enumerate(rows[1:]):
    name = r[ki].split("(")[0].replace("void ", "").replace("ofsv::", "")[-48:]
    us = float(r[vi].replace(",", "")) / 1000.0
    tot += us
    c = cls.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += us
    if "-v" in sys.argv:
        print(f"{i:3d} {name:50s} {us:9.1f} us  grid {r[gi]}")
print(f"total {tot:.1f} us over {len(rows) - 1} launches")
for k, (n, us) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    print(f"  {us:9.1f} us  {100 * us / tot:5.1f} %  x{n:<3d} {k}")
