"""Summarise an `ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list: one row per launch (time, warp instructions, DRAM bytes).  usage: launch_table2.py file.csv [min_us]"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
h = rows[0]
ki, mi, vi, gi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID"), h.index("Metric Unit")
d = OrderedDict()
for r in rows[1:]:
    e = d.setdefault(r[gi], {"name": r[ki].split("(")[0].replace("void ", "").replace("ofsv::", "")[:40]})
    v = float(r[vi].replace(",", ""))
    u = r[ui].lower()
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    e[r[mi]] = v
tot = sum(e["gpu__time_duration.sum"] for e in d.values())
for k, e in d.items():
    us = e["gpu__time_duration.sum"]
    if us >= thr:
        print(f"{k:>3} {e['name']:42s} {us:8.1f} us  inst {e.get('smsp__inst_executed.sum', 0) / 1e6:7.1f}M  "
              f"rd {e.get('dram__bytes_read.sum', 0) / 1e6:7.1f}MB wr {e.get('dram__bytes_write.sum', 0) / 1e6:7.1f}MB")
print(f"total {tot:.1f} us over {len(d)} launches")
