"""profiles/<out>.md from an ncu launch list of `python tools/run_train.py 1`: the launches between the last two adamw_kernel
launches = one full training step.   usage: python tools/summarize_train_launches.py gpurun_out/launches_X.csv profiles/Y.md"""
import csv, io, sys
from collections import OrderedDict

src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
rows = [r for r in csv.DictReader(io.StringIO("".join(lines))) if r.get("Metric Name") == "gpu__time_duration.sum"]
names = [r["Kernel Name"].split("(")[0] for r in rows]
idx = [i for i, n in enumerate(names) if "adamw_kernel" in n]
sel = rows[idx[-2] + 1: idx[-1] + 1]
agg = OrderedDict()
for r in sel:
    k = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
out = ["# ncu launch list — one full-size MIM training step (`python tools/run_train.py 1`, the launches between the last two `adamw_kernel`s)", "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).",
       f"{len(sel)} launches, {tot / 1e3:.2f} ms of kernel time (CUDA-event step time without ncu: see profiles/r01_bench_line.json).", "",
       "| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {n} | {t / 1e3:.3f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[:16]))
