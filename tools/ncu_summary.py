"""Summarise an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_*]` launch list per kernel:
launches, total / mean device time, share of the listed time, DRAM bytes per launch.  Usage:
    python tools/ncu_summary.py gpurun_out/launches.csv [skip_first_n_launches] > profiles/xxx.md"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = list(csv.reader(open(path)))
    hdr = None
    per_id = collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            e = per_id.setdefault(d["ID"], {"name": d["Kernel Name"], "grid": d["Grid Size"], "block": d["Block Size"]})
            e[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
            e["unit:" + d["Metric Name"]] = d["Metric Unit"]
    launches = list(per_id.values())[skip:]
    agg = collections.OrderedDict()
    for e in launches:
        name = e["name"].split("(")[0].replace("void ", "").replace("vocr::", "")
        t = e.get("gpu__time_duration.sum", 0.0)
        if e.get("unit:gpu__time_duration.sum", "ns") in ("us", "usecond"):
            t *= 1e3
        a = agg.setdefault(name, {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["ns"] += t
        a["rd"] += e.get("dram__bytes_read.sum", 0.0)
        a["wr"] += e.get("dram__bytes_write.sum", 0.0)
    total = sum(a["ns"] for a in agg.values())
    print("| kernel | launches | total ms | mean us | share | DRAM read MB/launch | DRAM write MB/launch |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print("| `%s` | %d | %.3f | %.1f | %.1f%% | %.2f | %.2f |" % (name, a["n"], a["ns"] / 1e6, a["ns"] / a["n"] / 1e3,
              100 * a["ns"] / total, a["rd"] / a["n"] / 1e6, a["wr"] / a["n"] / 1e6))
    print("\ntotal listed device time: %.3f ms over %d launches" % (total / 1e6, len(launches)))


if __name__ == "__main__":
    main()
