"""Share of GPU time per kernel from an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_shares.py file.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
d = OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui].lower(), 1.0)
    name = r[ki].split("(")[0].replace("void ", "").replace("ofsv::", "")[:60]
    e = d.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += v
tot = sum(e[1] for e in d.values())
print("# kernel, launches, total us, share")
for k, e in sorted(d.items(), key=lambda kv: -kv[1][1]):
    print(f"{k}, {e[0]}, {e[1]:.0f}, {100 * e[1] / tot:.1f} %")
print(f"# total {tot:.0f} us over {sum(e[0] for e in d.values())} launches")
