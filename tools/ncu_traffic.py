"""Mean DRAM traffic and duration per launch of the grouping kernels in an `ncu --set full` report of one
path-replay step -> profiles/ncu_traffic.json (read by bench.py for roofline.traffic) and a markdown table.

    python tools/ncu_traffic.py gpurun_out/step_group.ncu-rep profiles/ncu_traffic.json
"""
import csv
import json
import subprocess
import sys
from collections import defaultdict

rep, out_json = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]


def col(name):
    return hdr.index(name)


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


ENTRY = {"group_fwd_tma_kernel": "amc3d_group_points_ws", "group_bwd_tma_kernel": "amc3d_group_points_grad_ws"}
agg = defaultdict(lambda: [0, 0.0, 0.0])
ir, iw, it, ik = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum"), col("Kernel Name")
for r in rows[2:]:
    for k, e in ENTRY.items():
        if k in r[ik]:
            a = agg[e]
            a[0] += 1
            a[1] += to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
            t = float(r[it].replace(",", ""))
            a[2] += t * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[it], 1.0)
res = {e: round(a[1] / a[0]) for e, a in agg.items()}
if "amc3d_group_points_grad_ws" in res:      # the autograd path calls the overwrite variant of the same kernel
    res["amc3d_group_points_grad_ws_set"] = res["amc3d_group_points_grad_ws"]
json.dump(res, open(out_json, "w"), indent=1)
print("| entry point | launches | mean dram bytes / launch | mean kernel us |")
print("|---|---:|---:|---:|")
for e, a in agg.items():
    print(f"| `{e}` | {a[0]} | {a[1] / a[0] / 1e6:.1f} MB | {a[2] / a[0]:.1f} |")
