"""Per-instruction pc-sample listing of a kernel in an .ncu-rep (source page): shows every SASS line
whose share of samples exceeds a threshold, with cumulative share and executions per `unit`.

    python tools/ncu_hot.py rep.ncu-rep [min_pct] [exec_divisor]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iE, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) > iS and r[iS].isdigit()]
tot = sum(int(r[iS]) for r in body)
print("total samples", tot)
cum = 0
for k, r in enumerate(body):
    s = int(r[iS])
    cum += s
    if 100.0 * s / tot >= thr:
        st = sorted(((int(r[i]), h[6:]) for i, h in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:2]
        sts = " ".join(f"{n}:{100 * v / max(s, 1):.0f}%" for v, n in st)
        print(f"{k:4d} {100 * s / tot:5.1f}% cum {100 * cum / tot:5.1f}% exec {int(r[iE]) / div:9.2f}  {r[iSrc].strip()[:70]:70s} {sts}")
