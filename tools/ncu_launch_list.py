"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: per-kernel
launches per step and time per step (cold-cache, serialised launches: compare SHARES, not absolutes).
usage: python tools/ncu_launch_list.py launches.csv [steps_in_capture]"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None


def short(n):
    n = n.replace("(int)", "").replace("(bool)", "")
    m = re.search(r"gm::(\w+)(<[^(]*>)?", n)
    if m:
        t = m.group(2) or ""
        return m.group(1) + t.replace(" ", "")
    return "torch:" + n.split("(")[0][-60:]


agg = OrderedDict()
for r in rows:
    k = short(r[4])
    v = float(r[-1].replace(",", ""))
    unit = r[-2]
    us = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "nsecond": 1e-3, "msecond": 1e3}.get(unit, 1.0)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total us | share |")
print("|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / tot))
print("\ntotal %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
