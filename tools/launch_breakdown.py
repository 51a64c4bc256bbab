"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the LAST build pipeline (k_build_init ..
k_build_finalize) whose total exceeds `min_ns`, grouped by kernel.  usage: launch_breakdown.py file.csv [min_ns]"""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
min_ns = float(sys.argv[2]) if len(sys.argv) > 2 else 5e5
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
starts = [i for i, r in enumerate(data) if "k_build_init" in r[ix["Kernel Name"]]]
for s in reversed(starts):
    seg = []
    for r in data[s:]:
        seg.append(r)
        if "k_build_finalize" in r[ix["Kernel Name"]]:
            break
    tot = sum(float(r[ix["Metric Value"]]) for r in seg)
    if tot > min_ns:
        break
agg = OrderedDict()
for r in seg:
    k = r[ix["Kernel Name"]].split("(")[0][:60]
    agg[k] = agg.get(k, 0) + float(r[ix["Metric Value"]])
print("unit", seg[0][ix["Metric Unit"]], "total", tot, "launches", len(seg))
for k, v in agg.items():
    print(f"{v:14.0f}  {100 * v / tot:5.1f}%  {k}")
