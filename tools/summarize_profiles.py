#!/usr/bin/env python
"""Turn the raw ncu artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r1

Reads  gpurun_out/launches_<round>.csv   (ncu --metrics gpu__time_duration.sum launch list of bench.py)
       gpurun_out/<round>_<name>.ncu-rep (ncu --set full captures of single launches)
       gpurun_out/bench_<round>.json     (the bench line of the same code)
Writes profiles/<round>_launches.csv, <round>_launch_summary.md, <round>_ncu_<name>.txt, <round>_traffic.json,
       <round>_bench.json
"""
import collections
import csv
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
src, dst = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(dst, exist_ok=True)

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "lts__t_bytes.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def top_stalls(rep, n=12):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ci, cs, ce = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = []
    for r in rows[2:]:
        try:
            data.append((float(r[cs]), r[ci].strip()[:100], r[ce]))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1.0
    return [(100 * s / tot, t, e) for s, t, e in sorted(data, reverse=True)[:n]]


traffic = {}
for rep in sorted(glob.glob(os.path.join(src, f"{rnd}_*.ncu-rep"))):
    name = os.path.basename(rep)[len(rnd) + 1:-8]
    m = raw_metrics(rep)
    lines = [f"# ncu --set full --clock-control none --import-source on, one launch of `{m['Kernel Name'][0][:120]}`",
             f"# captured inside `python bench.py --steps 2 --warmup 3` on a B200; round {rnd}", ""]
    for k in KEEP:
        if k in m:
            lines.append(f"{k:75s} {m[k][0]:>18s} {m[k][1]}")
    stall = [(k, v) for k, v in m.items() if "issue_stalled" in k and "per_issue_active" in k]
    lines.append("")
    lines.append("# warp stall reasons (warps per issue-active cycle)")
    for k, (v, u) in sorted(stall, key=lambda kv: -float(kv[1][0].replace(",", "") or 0))[:8]:
        lines.append(f"{k:90s} {v}")
    lines.append("")
    lines.append("# hottest SASS instructions by stall samples (% of samples, executed count)")
    try:
        for pct, text, ex in top_stalls(rep):
            lines.append(f"{pct:5.1f}%  exec {ex:>10s}  {text}")
    except Exception as e:
        lines.append(f"(source page unavailable: {e})")
    open(os.path.join(dst, f"{rnd}_ncu_{name}.txt"), "w").write("\n".join(lines) + "\n")
    rd = float(m["dram__bytes_read.sum"][0].replace(",", "")) * UNIT.get(m["dram__bytes_read.sum"][1], 1)
    wr = float(m["dram__bytes_write.sum"][0].replace(",", "")) * UNIT.get(m["dram__bytes_write.sum"][1], 1)
    traffic[name] = {"kernel": m["Kernel Name"][0].split("(")[0], "dram_bytes_per_launch": rd + wr,
                     "duration_us_under_ncu": m["gpu__time_duration.sum"][0] + " " + m["gpu__time_duration.sum"][1],
                     "tensor_pipe_pct": m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", ("", ""))[0],
                     "dram_pct": m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ("", ""))[0]}
    print("wrote", name)
json.dump(traffic, open(os.path.join(dst, f"{rnd}_traffic.json"), "w"), indent=1)

lc = os.path.join(src, f"launches_{rnd}.csv")
if os.path.exists(lc):
    shutil.copyfile(lc, os.path.join(dst, f"{rnd}_launches.csv"))
    lines = [l for l in open(lc) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for row in r:
        name = row[ki].split("(")[0].replace("void ", "")
        v = float(row[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0}.get(row[ui], 1e-6)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    md = [f"# ncu launch list, round {rnd}", "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cra5 -s 900 -c 300 python bench.py --steps 2 "
          "--warmup 3` (cold-cache, serialised: compare SHARES with the bench's own CUDA-event table, not absolutes).",
          "", f"{n} launches, {tot:.2f} ms summed", "", "| kernel | launches | ms | share |", "|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f} % |")
    open(os.path.join(dst, f"{rnd}_launch_summary.md"), "w").write("\n".join(md) + "\n")
    print("wrote launch summary")
bj = os.path.join(src, f"bench_{rnd}.json")
if os.path.exists(bj):
    line = [l for l in open(bj) if l.startswith("{")][-1]
    json.dump(json.loads(line), open(os.path.join(dst, f"{rnd}_bench.json"), "w"), indent=1)
    print("wrote bench json")
