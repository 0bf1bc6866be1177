"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list by kernel.
Usage: python tools/ncu_launches.py launches.csv "<command that was profiled>" > summary.txt"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    name, metric, unit, val = r[ix["Kernel Name"]], r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    a = agg[name]
    if metric == "gpu__time_duration.sum":
        a[0] += val / (1e3 if unit in ("ns", "nsecond") else 1.0)
        a[1] += 1
    else:
        mb = val / {"byte": 1e6, "Kbyte": 1e3, "Mbyte": 1.0, "Gbyte": 1e-3}.get(unit, 1e6)
        a[2 if "read" in metric else 3] += mb
tot = sum(a[0] for a in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print(f"total {tot:.1f} us over {sum(a[1] for a in agg.values())} launches (raw list: {sys.argv[1].split('/')[-1]})\n")
print("  us total calls  share  us/call  dram rd MB/call  dram wr MB/call  kernel")
for name, (us, n, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:10.1f} {n:5d} {100 * us / tot:5.1f}% {us / n:8.1f} {rd / n:16.1f} {wr / n:16.1f}  {name[:120]}")
