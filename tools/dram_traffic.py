"""dram__bytes_read.sum + dram__bytes_write.sum per launch of the step's kernels, from an `ncu --set full` report, as
the JSON bench.py quotes for `roofline.traffic` -- stamped with the sha256 of the csrc/ sources the capture was taken
from (bench.py reports the number only while that digest still matches the tree).
Usage: python tools/dram_traffic.py report.ncu-rep out.json "<how the capture was taken>" """
import csv
import glob
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"march_rays_kernel": "march_rays", "nerf_fused_forward_kernel": "nerf_fused_forward", "integrate_loss_fused_kernel": "integrate_loss_fused (replaces the three above in the step)",
        "nerf_mlp_backward_umma_kernel<(bool)0>": "nerf_mlp_backward", "nerf_mlp_backward_umma_kernel<(bool)1>": "nerf_mlp_backward_scatter (replaces the two above in the step)",
        "nerf_mlp_backward_umma_kernel<0>": "nerf_mlp_backward", "nerf_mlp_backward_umma_kernel<1>": "nerf_mlp_backward_scatter (replaces the two above in the step)", "hashgrid_a1_backward_kernel": "hashgrid_a1_backward", "adam_kernel": "adam_step"}
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in data:
    name = r[ix["Kernel Name"]]
    for sub, key in KEYS.items():
        if sub in name:
            b = sum(float(r[ix[m]].replace(",", "")) * scale.get(units[ix[m]], 1) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            acc.setdefault(key, []).append(b)
h = hashlib.sha256()
for f in sorted(glob.glob(os.path.join(ROOT, "jaxngp_b200", "csrc", "*.cu*"))):
    h.update(open(f, "rb").read())
out = {k: int(sum(v) / len(v)) for k, v in acc.items()}
out["csrc_sha256"] = h.hexdigest()
out["file"] = os.path.basename(sys.argv[2])
out["_source"] = sys.argv[3] if len(sys.argv) > 3 else ""
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
