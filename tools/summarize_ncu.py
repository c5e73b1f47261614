"""Reads an `ncu --set full` report (here, no GPU needed) and writes tracked summaries under profiles/:
    python tools/summarize_ncu.py gpurun_out/prof_r02_kernels.ncu-rep r02 [gpurun_out/ncu_targets_order.json]
-> profiles/<tag>_ncu_kernels.md (one row per captured launch: duration, tensor-pipe %, XU %, issue-active %, DRAM bytes and %, L2
   throughput, registers) and profiles/<tag>_ncu_traffic.json (dram bytes per launch per kernel; bench.py's `roofline.traffic`)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
order = json.load(open(sys.argv[3])) if len(sys.argv) > 3 and os.path.exists(sys.argv[3]) else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rr[0], rr[1], rr[2:]
col = {h: i for i, h in enumerate(hdr)}


def num(row, name):
    i = col.get(name)
    if i is None or row[i] in ("", "n/a"):
        return None
    try:
        return float(row[i].replace(",", ""))
    except ValueError:
        return None


def to_bytes(row, name):
    v = num(row, name)
    if v is None:
        return None
    u = units[col[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def to_us(row, name):
    v = num(row, name)
    u = units[col[name]].lower()
    return None if v is None else v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(u, 1)


M = dict(tensor="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", tensor_mem="sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
         xu="sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", fma="sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         alu="sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", issue="smsp__issue_active.avg.pct_of_peak_sustained_active",
         dram_pct="gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", lts_pct="lts__throughput.avg.pct_of_peak_sustained_elapsed",
         sm_pct="sm__throughput.avg.pct_of_peak_sustained_elapsed", regs="launch__registers_per_thread", grid="launch__grid_size", block="launch__block_size")
if M["tensor"] not in col:  # metric name differs between ncu versions: take any tensor-pipe "cycles active" percentage
    for h in hdr:
        if "pipe_tensor" in h and "pct_of_peak_sustained_active" in h:
            M["tensor"] = h
            break
rows, traffic = [], {}
for i, r in enumerate(vals):
    name = r[col["Kernel Name"]].split("(")[0]
    o = order[i] if i < len(order) and order[i]["kernel"].split("_kernel")[0] in name else {}
    dur = to_us(r, "gpu__time_duration.sum")
    rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
    d = dict(i=i, kernel=name, shape=o.get("shape", ""), us=dur, dram=(rd or 0) + (wr or 0), rd=rd, wr=wr, alg_bytes=o.get("algorithmic_bytes"), alg_flops=o.get("algorithmic_flops"))
    for k, mname in M.items():
        d[k] = num(r, mname)
    rows.append(d)
    traffic[f"{name} [{o.get('shape', i)}]"] = dict(dram_bytes_per_launch=d["dram"], dram_read=rd, dram_write=wr, shape=o.get("shape", ""), duration_us_under_ncu=dur,
                                                      algorithmic_bytes=o.get("algorithmic_bytes"), tensor_pipe_pct=d.get("tensor"))
out = os.path.join(ROOT, "profiles")
f2 = lambda v, fmt="{:.1f}": "-" if v is None else fmt.format(v)
with open(os.path.join(out, f"{tag}_ncu_kernels.md"), "w") as f:
    f.write(f"# `ncu --set full --clock-control none` of every hot kernel at its benchmark shape (`python tools/ncu_targets.py`), tag {tag}\n\n")
    f.write("One launch each, cold caches (ncu flushes between replays), serialised: durations are NOT bench numbers (CUDA-event times are in the\n"
            "bench line / profiles/*gpu_check*); read the pipe percentages, DRAM bytes against the algorithmic bytes, registers.\n\n")
    f.write("| # | kernel | shape | us (ncu) | tensor pipe % | tensor mem % | XU % | FMA % | ALU % | issue % | DRAM MB | algorithmic MB | DRAM % | L2 % | regs | grid x block |\n")
    f.write("|---:|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
    for d in rows:
        f.write(f"| {d['i']} | `{d['kernel']}` | {d['shape']} | {f2(d['us'])} | {f2(d['tensor'])} | {f2(d['tensor_mem'])} | {f2(d['xu'])} | {f2(d['fma'])} | {f2(d['alu'])} | "
                f"{f2(d['issue'])} | {f2(d['dram'] / 1e6)} | {f2(None if d['alg_bytes'] is None else d['alg_bytes'] / 1e6)} | {f2(d['dram_pct'])} | {f2(d['lts_pct'])} | "
                f"{f2(d['regs'], '{:.0f}')} | {f2(d['grid'], '{:.0f}')} x {f2(d['block'], '{:.0f}')} |\n")
json.dump(dict(source=os.path.basename(rep), how="ncu --set full --clock-control none, one launch per kernel, dram__bytes_read.sum + dram__bytes_write.sum",
               kernels=traffic), open(os.path.join(out, f"{tag}_ncu_traffic.json"), "w"), indent=1)
print(open(os.path.join(out, f"{tag}_ncu_kernels.md")).read())
