"""Summarise an ncu --set full --import-source on report: headline metrics, stall mix of the hot loop, hottest SASS.
usage: python tools/ncu_hot.py report.ncu-rep [min_exec]"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
min_exec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, v = r[0], r[-1]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
d = dict(zip(h, v))
for k in want:
    if k in d: print(f"{k:80s} {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, x in enumerate(rows) if "Source" in x and "Address" in x)
hdr, data = rows[hi], rows[hi + 1:]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
sel = [x for x in data if x[iex].isdigit() and int(x[iex]) >= min_exec]
tot = sum(int(x[isamp] or 0) for x in sel)
print(f"\n# {len(sel)} SASS instructions with >= {min_exec} executions, {tot} samples")
mix = collections.Counter()
for x in sel:
    for i in stall_cols:
        if x[i].isdigit(): mix[hdr[i]] += int(x[i])
print("stall mix:", ", ".join(f"{k[6:]} {100 * n / max(tot, 1):.1f}%" for k, n in mix.most_common(10)))
ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[ia].strip()).split()[0].split(".")[0] for x in sel)
print("opcode mix:", ", ".join(f"{k} {n}" for k, n in ops.most_common(14)))
print("hottest:")
for x in sorted(sel, key=lambda x: -int(x[isamp] or 0))[:25]:
    st = max(((int(x[i]) if x[i].isdigit() else 0, hdr[i][6:]) for i in stall_cols))
    print(f" {100 * int(x[isamp] or 0) / max(tot, 1):5.1f}%  exec {x[iex]:>9}  {st[1]:14s} {x[ia][:80]}")
