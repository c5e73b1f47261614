"""Summarise ncu output brought back in gpurun_out/ into tracked files under profiles/.
usage: python tools/summarize_profile.py <tag>   (reads gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep)"""
import csv
import io
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list -------------------------------------------------------------------------
p = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
lines = [l for l in open(p) if l.startswith('"')]
rows = list(csv.DictReader(io.StringIO("".join(lines))))
agg = OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v_us
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
    f.write(f"# ncu launch list `{tag}` — `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mim`\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 600` (cold-cache, serialised: compare SHARES, not absolutes).\n")
    f.write(f"{len(rows)} launches captured (warm-up + device-timed + e2e passes), {tot/1e3:.2f} ms of kernel time.\n\n")
    f.write("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {t/1e3:.3f} | {t/n:.1f} | {100*t/tot:.1f}% |\n")
print(open(os.path.join(out, f"{tag}_launches.md")).read())

# ---- full capture ------------------------------------------------------------------------
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rr[0], rr[1], rr[2:]
    want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
    with open(os.path.join(out, f"{tag}_ncu_full.md"), "w") as f:
        f.write(f"# ncu --set full capture `{tag}` (`--clock-control none --import-source on`), {len(vals)} launch(es)\n\n")
        f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(vals))) + " |\n|---|---|" + "---|" * len(vals) + "\n")
        for w in want:
            for i, h in enumerate(hdr):
                if h == w:
                    f.write(f"| {h} | {units[i]} | " + " | ".join(v[i][:90] for v in vals) + " |\n")
    print(open(os.path.join(out, f"{tag}_ncu_full.md")).read())
