"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/ncu_summary.py gpurun_out/launches.csv [out.md]"""
import csv, sys, re, collections

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    us = val * scale
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1; a[1] += us; a[2] = max(a[2], us)
    total += us
out = ["| kernel | launches | total ms | share | avg us | max us |", "|---|---:|---:|---:|---:|---:|"]
for name, (n, t, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{name[:90]}` | {n} | {t/1e3:.2f} | {100*t/total:.1f}% | {t/n:.2f} | {mx:.1f} |")
out.append(f"| **total** | {sum(a[0] for a in agg.values())} | {total/1e3:.2f} | 100% | | |")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
