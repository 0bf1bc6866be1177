"""Per-T table of the C4 sweep under ncu: python tools/ncu_hashenc.py launches.csv > table.txt
Input: `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,
lts__t_sector_hit_rate.pct -k regex:hashgrid` of `HASHENC_SWEEP_ONCE=1 python tools/hashenc_sweep.py`, which launches, per
table size T: forward + backward single-pass (point-major, NGP_B200_HG_LPG=16), then forward + backward with the launcher's
own policy (level-major passes beyond L2)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["ID"]] == "ID":
        continue
    launches.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]].split("(")[0].split("::")[-1][:44], "grid": r[ix["Grid Size"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
unit = {r[ix["Metric Name"]]: r[ix["Metric Unit"]] for r in rows[1:] if r[ix["ID"]] != "ID"}
print(f"{'id':>4s} {'kernel':44s} {'grid':>16s} {'us':>9s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'L2 MB':>10s} {'L2 hit %':>8s}")
scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
tscale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
for i, m in launches.items():
    def mb(k):
        return m.get(k, 0.0) * scale.get(unit.get(k, "byte"), 1e-6)
    us = m.get("gpu__time_duration.sum", 0.0) * tscale.get(unit.get("gpu__time_duration.sum", "ns"), 1e-3)
    print(f"{i:4d} {m['kernel']:44s} {m['grid']:>16s} {us:9.1f} {mb('dram__bytes_read.sum'):11.1f} {mb('dram__bytes_write.sum'):11.1f} "
          f"{mb('lts__t_bytes.sum'):10.1f} {m.get('lts__t_sector_hit_rate.pct', 0.0):8.1f}")
