"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.

    python tools/summarize_launches.py gpurun_out/launches_r01.csv [steps] > profiles/r01_launches.md

The per-launch times ncu reports are cold-cache and serialised, so only the SHARES are meaningful
(B200_PROFILING.md); bench.py's live CUDA-event shares must agree with these.
"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    if name.startswith("at::") or "at::native" in name:
        m = re.search(r"(\w+_kernel\w*|\w+Kernel\w*)", name)
        return "torch:" + (m.group(1) if m else name[:40])
    name = re.sub(r"\(.*$", "", name)
    return name.replace("amc3d::", "")


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", "")) / 1e3,
                         r["Grid Size"], r["Block Size"]))
    agg = defaultdict(lambda: [0, 0.0])
    for name, us, _, _ in rows:
        agg[name][0] += 1
        agg[name][1] += us
    total = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if not k.startswith("torch:"))
    print(f"source: {path}  ({len(rows)} launches, {steps} profiled passes incl. warm-up; times are ncu-serialised, "
          f"cold-cache: compare shares only)\n")
    print(f"total {total / 1e3:.2f} ms, of which this library's kernels {ours / 1e3:.2f} ms "
          f"({100 * ours / total:.1f} %), torch glue {100 * (total - ours) / total:.1f} %\n")
    print("| kernel | launches | total us | share | mean us |")
    print("|---|---:|---:|---:|---:|")
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {cnt} | {us:.1f} | {100 * us / total:.2f} % | {us / cnt:.2f} |")


if __name__ == "__main__":
    main()
