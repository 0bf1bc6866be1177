"""Top stall sites of one kernel from an ncu report's SASS page:
python tools/ncu_hotspots.py report.ncu-rep <kernel regex> [top N]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
ends = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
rows = rows[ends[0]:ends[1]] if len(ends) > 1 else rows[ends[0]:]
print(rows[0][1][:100])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot)
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    s = {k: int(r[ix[k]]) for k in stalls if int(r[ix[k]])}
    main = sorted(s.items(), key=lambda kv: -kv[1])[:3]
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {r[ix['Source']].strip()[:70]:70s} {main}")
