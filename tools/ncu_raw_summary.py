"""Turn an `ncu -i X.ncu-rep --page raw --csv` dump into the markdown table kept under profiles/.
usage: python tools/ncu_raw_summary.py raw.csv > table.md"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
launches = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
unit = dict(zip(hdr, units))
names = [d.get("Kernel Name", "?").split("(")[0] for d in launches]
print("| metric | unit | " + " | ".join(names) + " |")
print("|---|---|" + "---|" * len(launches))
for w in WANT:
    if w in hdr:
        print(f"| `{w}` | {unit[w]} | " + " | ".join(d[w] for d in launches) + " |")
stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
def val(d, h):
    try:
        return float(d[h].replace(",", ""))
    except ValueError:
        return 0.0
stalls.sort(key=lambda h: -val(launches[-1], h))
for h in stalls[:9]:
    print(f"| `{h}` | {unit[h]} | " + " | ".join(f"{val(d, h):.3f}" for d in launches) + " |")
