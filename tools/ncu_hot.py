"""Developer tool: condensed view of an `ncu --page source --csv` dump: stall totals and contiguous regions of equal execution count.
usage: python tools/ncu_hot.py dump.csv [min_exec]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
base = int(data[0][0], 16)
tot_i = sum(f(r, 'Instructions Executed') for r in data)
tot_s = sum(f(r, '# Samples') for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print(f"warp inst {tot_i/1e6:.1f}M  thread inst {sum(f(r,'Thread Instructions Executed') for r in data)/1e9:.2f}G  samples {tot_s:.0f}")
print("  ".join(f"{s[6:]} {100*sum(f(r,s) for r in data)/tot_s:.1f}%" for s in sorted(stalls, key=lambda s: -sum(f(r, s) for r in data))[:9]))
# regions: consecutive instructions whose exec count differs < 2 % and lanes differ < 1
reg = []
for r in data:
    ex, ln = f(r, 'Instructions Executed'), f(r, 'Avg. Threads Executed')
    if reg and abs(reg[-1]['ex'] - ex) <= 0.02 * max(ex, 1) and abs(reg[-1]['ln'] - ln) < 1.0:
        g = reg[-1]
    else:
        g = dict(addr=int(r[0], 16) - base, ex=ex, ln=ln, n=0, inst=0.0, smp=0.0, first=r[1].strip(), st={})
        reg.append(g)
    g['n'] += 1; g['inst'] += ex; g['smp'] += f(r, '# Samples')
    for s in stalls: g['st'][s] = g['st'].get(s, 0) + f(r, s)
mn = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
print(f"{'addr':>6} {'n':>4} {'exec':>8} {'%inst':>6} {'lanes':>5} {'%smp':>5}  top stall / first instruction")
for g in reg:
    if 100 * g['inst'] / tot_i < mn and 100 * g['smp'] / tot_s < mn: continue
    top = max(g['st'], key=lambda s: g['st'][s])
    print(f"{g['addr']:6x} {g['n']:4d} {g['ex']/1e6:7.2f}M {100*g['inst']/tot_i:6.2f} {g['ln']:5.1f} {100*g['smp']/tot_s:5.1f}  {top[6:]:14s} {g['first'][:60]}")
