#!/usr/bin/env python
"""Summarise every kernel launch of an ncu report (--set full [--import-source on]) into small text files + one JSON.
Runs ON the GPU box right after the capture, so that only the summaries travel back (reports are tens of MB).

    python tools/ncu_summarize.py <report.ncu-rep> <out_dir> <prefix>
"""
import collections, csv, io, json, os, re, subprocess, sys

rep, out_dir, prefix = sys.argv[1], sys.argv[2], sys.argv[3]
os.makedirs(out_dir, exist_ok=True)
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return 0.0


raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, launches = rows[0], rows[1], rows[2:]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
heads = [i for i, x in enumerate(srows) if "Source" in x and "Address" in x]
summary = {}
for li, vals in enumerate(launches):
    m = {h: (vals[i], units[i]) for i, h in enumerate(hdr) if i < len(vals)}
    kname = m["Kernel Name"][0]
    short = re.sub(r"[^A-Za-z0-9_]+", "_", kname.split("(")[0].replace("void ", "").replace("cra5::", "").replace("<unnamed>::", ""))[:48].strip("_")
    name = f"{prefix}_{li:02d}_{short}"
    lines = [f"# ncu --set full --clock-control none --import-source on: launch {li} of `{kname[:150]}`",
             f"# captured inside `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-profile --no-e2e` (8 frames per call) on a B200 (report {os.path.basename(rep)})", ""]
    for k in KEEP:
        if k in m:
            lines.append(f"{k:78s} {m[k][0]:>18s} {m[k][1]}")
    stall = [(k, v) for k, v in m.items() if "issue_stalled" in k and "per_issue_active" in k]
    lines += ["", "# warp stall reasons (warps per issue-active cycle)"]
    for k, (v, u) in sorted(stall, key=lambda kv: -num(kv[1][0]))[:8]:
        lines.append(f"{k:90s} {v}")
    if li < len(heads):
        h = srows[heads[li]]
        end = heads[li + 1] - 1 if li + 1 < len(heads) else len(srows)
        data = [x for x in srows[heads[li] + 1:end] if len(x) == len(h)]
        ci, cs, ce = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        tot = sum(num(x[cs]) for x in data) or 1.0
        lines += ["", "# hottest SASS instructions by stall samples (% of samples, executed count)"]
        for x in sorted(data, key=lambda x: -num(x[cs]))[:12]:
            lines.append(f"{100 * num(x[cs]) / tot:5.1f}%  exec {x[ce]:>10s}  {x[ci].strip()[:100]}")
    open(os.path.join(out_dir, name + ".txt"), "w").write("\n".join(lines) + "\n")
    rd = num(m["dram__bytes_read.sum"][0]) * UNIT.get(m["dram__bytes_read.sum"][1], 1)
    wr = num(m["dram__bytes_write.sum"][0]) * UNIT.get(m["dram__bytes_write.sum"][1], 1)
    summary[name] = {"kernel": kname.split("(")[0], "grid": m.get("launch__grid_size", ("", ""))[0],
                     "dram_bytes_per_launch": rd + wr,
                     "duration_us_under_ncu": m["gpu__time_duration.sum"][0] + " " + m["gpu__time_duration.sum"][1],
                     "tensor_pipe_pct": m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", ("", ""))[0],
                     "dram_pct": m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ("", ""))[0]}
json.dump(summary, open(os.path.join(out_dir, prefix + "_summary.json"), "w"), indent=1)
print("summarised", len(launches), "launches of", rep)
