"""Per-kernel summary of an `ncu --set full` report -> markdown (profiles/r02_kernels_ncu.md).

    python tools/ncu_kernel_table.py gpurun_out/r02_step.ncu-rep > profiles/r02_kernels_ncu.md

Per kernel name (template arguments kept): launches, total and mean duration, issue-slot utilisation, achieved
occupancy, DRAM and L2 bytes per launch, registers, and the top stall reasons (ratios per issued instruction)."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
if rep.endswith(".csv"):                       # `ncu -i report --page raw --csv` saved on the GPU box (the report itself is large)
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(out.splitlines()) if r]
while rows and "Kernel Name" not in rows[0]:
    rows.pop(0)
hdr, units = rows[0], rows[1]
ci = {n: i for i, n in enumerate(hdr)}


def num(r, name):
    try:
        return float(r[ci[name]].replace(",", ""))
    except Exception:
        return 0.0


def to_bytes(r, name):
    if name not in ci:
        return 0.0
    u = units[ci[name]]
    return num(r, name) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(r, name):
    if name not in ci:
        return 0.0
    u = units[ci[name]]
    return num(r, name) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)


def short(n):
    n = re.sub(r"^void\s+", "", n).replace("amc3d::", "")
    return re.sub(r"\(.*$", "", n)


stall_cols = [c for c in hdr if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")
              and "not_issued" not in c]
agg = defaultdict(lambda: defaultdict(float))
for r in rows[2:]:
    k = short(r[ci["Kernel Name"]])
    a = agg[k]
    a["n"] += 1
    a["us"] += to_us(r, "gpu__time_duration.sum")
    a["issue"] += num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    a["occ"] += num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    a["dram"] += to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")
    a["l2"] += to_bytes(r, "lts__t_bytes.sum") or 32.0 * num(r, "lts__t_sectors.sum")
    a["regs"] = num(r, "launch__registers_per_thread")
    a["fp32"] += max(num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                     num(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"))
    a["tensor"] += num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    a["dramp"] += num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    for c in stall_cols:
        a["st_" + c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] += num(r, c)
print("| kernel | launches | total us | mean us | issue slots busy % | warps active % | FMA pipe % | tensor pipe % | DRAM % of peak | DRAM MB / launch | L2 MB / launch | regs | top stalls (per issued instruction) |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = a["n"]
    stalls = sorted(((v / n, s[3:]) for s, v in a.items() if s.startswith("st_")), reverse=True)[:3]
    st = ", ".join(f"{name} {v:.2f}" for v, name in stalls)
    print(f"| `{k}` | {int(n)} | {a['us']:.1f} | {a['us'] / n:.1f} | {a['issue'] / n:.1f} | {a['occ'] / n:.1f} | {a['fp32'] / n:.1f} | {a['tensor'] / n:.1f} | {a['dramp'] / n:.1f} | "
          f"{a['dram'] / n / 1e6:.1f} | {a['l2'] / n / 1e6:.1f} | {int(a['regs'])} | {st} |")
