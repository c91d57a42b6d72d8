"""`ncu -i X.ncu-rep --page raw --csv` -> one markdown row per captured launch with the counters DESIGN.md argues from.
    python tools/ncu_full_summary.py gpurun_out/r02_top_kernels_raw.csv > profiles/r02_top_kernels_ncu.md"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, scale=1.0, fmt="%.1f"):
    i = col.get(name)
    if i is None or r[i] in ("", "n/a"): return ""
    try: return fmt % (float(r[i].replace(",", "")) * scale)
    except ValueError: return r[i]
unit_t = rows[1][col["gpu__time_duration.sum"]]
ts = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit_t, 1.0)
ub = rows[1][col["dram__bytes_read.sum"]]
bs = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(ub, 1e-6)
print("| kernel | grid | time us | tensor pipe active % | SM throughput % | DRAM throughput % | DRAM read MB | DRAM write MB | L2 hit % | warps active % | regs/thread |")
print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("vocr::", "")
    print("| `%s` | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
        name, r[col["launch__grid_size"]], g(r, "gpu__time_duration.sum", ts), g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"), g(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
        g(r, "dram__bytes_read.sum", bs), g(r, "dram__bytes_write.sum", bs), g(r, "lts__t_sector_hit_rate.pct"),
        g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), g(r, "launch__registers_per_thread", 1.0, "%d")))
