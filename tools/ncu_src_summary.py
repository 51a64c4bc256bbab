"""Summarise an `ncu --page source --csv` dump: per basic region, instructions executed, lane utilisation, stall samples.
Usage: python tools/ncu_src_summary.py dump.csv [nregions]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        data.append((int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]]), r[ix["Source"]],
                     int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0), int(r[ix["Thread Instructions Executed"]] or 0), r))
    except ValueError:
        continue
seen = set()
dd = []
for d in data:
    if d[0] in seen:
        continue
    seen.add(d[0])
    dd.append(d)
data = dd
base = data[0][0]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
tot_inst = sum(d[3] for d in data)
tot_thr = sum(d[4] for d in data)
tot_smp = sum(d[2] for d in data)
print(f"total warp-inst {tot_inst:,}  thread-inst {tot_thr:,}  avg lanes {tot_thr / max(tot_inst, 1):.2f}  samples {tot_smp:,}")
# regions: split where the execution count changes by > 2% between consecutive instructions
regions = []
cur = [data[0]]
for d in data[1:]:
    prev = cur[-1]
    if prev[3] == 0 and d[3] == 0:
        cur.append(d)
        continue
    if abs(d[3] - prev[3]) > 0.02 * max(d[3], prev[3], 1) or abs(d[4] - prev[4]) > 0.03 * max(d[4], prev[4], 1):
        regions.append(cur)
        cur = [d]
    else:
        cur.append(d)
regions.append(cur)
out = []
for reg in regions:
    ni = sum(d[3] for d in reg)
    nt = sum(d[4] for d in reg)
    ns = sum(d[2] for d in reg)
    out.append((ni, nt, ns, reg))
print(f"{'addr':>8} {'n':>4} {'exec/inst':>12} {'%inst':>6} {'lanes':>6} {'%samples':>8}  first instruction")
for ni, nt, ns, reg in out:
    if ni < thr * tot_inst and ns < thr * tot_smp:
        continue
    print(f"{reg[0][0] - base:8x} {len(reg):4d} {reg[0][3]:12,} {100 * ni / tot_inst:6.2f} {nt / max(ni, 1):6.2f} {100 * ns / max(tot_smp, 1):8.2f}  {reg[0][1][:70]}")
