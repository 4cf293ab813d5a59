#!/usr/bin/env python
"""One-call profiling recipe for a round, run ON the GPU box:

    gpurun --timeout 2400 -- python tools/profile_r2.py r2

  1. `python bench.py` (default flags: the line the driver will see) -> gpurun_out/prof/bench_<round>.json
  2. ncu launch list of one timed step of the same command (gpu__time_duration.sum, --clock-control none: cold-cache
     and serialised, so compare SHARES with the bench's own CUDA-event table, not absolutes)
  3. ncu --set full --import-source on captures of single launches picked from that list by position -- the four GEMMs
     of one global trunk block, patch-embed and un-patchify, the three window shapes and the global attention, LayerNorm,
     frame_to_patches, the fused quantise + index kernel, rANS encode / decode -- summarised on the box
     (tools/ncu_summarize.py: reports are tens of MB, only the text travels back)
  4. per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass): UTCHMMA / LDTM / STTM / UTMALDG ...

Copy what should be judged into profiles/ with `python tools/profile_r2.py --collect r2` (run in the build container).
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "prof")
BENCH = [sys.executable, os.path.join(ROOT, "bench.py")]
PROF_ARGS = ["--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-kernel-profile", "--no-e2e"]
KREGEX = ("regex:attn_|gemm_tc|layernorm|rans_|frame_to|gc_quant|eb_quant|scan_len|compact_|container_|im2col|"
          "transpose_cast|cast_bf16|word_to|split_rows")


def sh(cmd, timeout, **kw):
    try:
        return subprocess.run(cmd, timeout=timeout, capture_output=True, text=True, **kw)
    except subprocess.TimeoutExpired:
        print("TIMEOUT", " ".join(cmd[:6]), flush=True)
        return None


def launch_list(rnd):
    path = os.path.join(OUT, f"launches_{rnd}.csv")
    sh(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-k", KREGEX, "-c", "6000", "--csv",
        "--log-file", path] + BENCH + PROF_ARGS, 900)
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki = hdr.index("Kernel Name")
    return [r[ki] for r in rows[1:]]


def positions(names, pattern):
    """indexes (among the launches matching `pattern`) of the launches of the LAST of the four steps"""
    m = [i for i, n in enumerate(names) if re.search(pattern, n)]
    per_step = len(m) // 4
    return per_step, 3 * per_step      # launches per step, first launch of the timed step (pattern-relative)


def capture(rnd, tag, kregex, skip, count):
    rep = f"/tmp/cap_{tag}"
    r = sh(["ncu", "--set", "full", "--import-source", "on", "--clock-control", "none", "-k", "regex:" + kregex, "-s", str(skip),
            "-c", str(count), "-o", rep, "-f"] + BENCH + PROF_ARGS, 900)
    if r is None or not os.path.exists(rep + ".ncu-rep"):
        print("capture failed:", tag, (r.stderr[-500:] if r else ""), flush=True)
        return
    sh([sys.executable, os.path.join(ROOT, "tools", "ncu_summarize.py"), rep + ".ncu-rep", OUT, f"{rnd}_{tag}"], 600)
    os.remove(rep + ".ncu-rep")


def sass_histogram(rnd):
    lib = os.path.join(ROOT, "cra5_b200", "lib", "libcra5b200.so")
    r = sh(["cuobjdump", "-sass", lib], 600)
    if r is None:
        return
    hist, cur = collections.OrderedDict(), None
    for line in r.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = hist.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            cur[m.group(1).split(".")[0] + ("." + ".".join(m.group(1).split(".")[1:3]) if m.group(1).startswith(("UTC", "UTMA", "LDTM", "STTM")) else "")] += 1
    keys = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "MUFU", "FFMA2", "FADD2")
    with open(os.path.join(OUT, f"{rnd}_sass_opcodes.txt"), "w") as f:
        f.write("# cuobjdump -sass cra5_b200/lib/libcra5b200.so: per-kernel opcode counts (tensor-core / TMEM / TMA families\n"
                "# first: UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor load)\n\n")
        for fn, c in hist.items():
            total = sum(c.values())
            if total < 50:
                continue
            demangled = sh(["cu++filt", fn], 20)
            name = (demangled.stdout.strip() if demangled and demangled.stdout.strip() else fn)[:140]
            fam = collections.Counter()
            for op, n in c.items():
                for k in keys:
                    if op.startswith(k):
                        fam[op] += n
            f.write(f"{name}\n    {total} instructions; " + ", ".join(f"{k} {v}" for k, v in sorted(fam.items())) + "\n")
            f.write("    top: " + ", ".join(f"{k} {v}" for k, v in c.most_common(10)) + "\n\n")


def main(rnd):
    os.makedirs(OUT, exist_ok=True)
    partial = any(a.startswith("--only=") for a in sys.argv)
    if not partial:
        with open(os.path.join(OUT, f"bench_{rnd}.json"), "w") as f:
            r = subprocess.run(BENCH, stdout=f, stderr=subprocess.PIPE, text=True, timeout=1200)
        print("bench rc", r.returncode, r.stderr[-300:], flush=True)
    try:
        d = json.loads([l for l in open(os.path.join(OUT, f"bench_{rnd}.json")) if l.startswith("{")][-1])
        print("value", d["value"], "e2e", (d.get("e2e") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), flush=True)
    except Exception as e:
        print("bench line unreadable:", e, flush=True)
    names = launch_list(rnd)
    print(len(names), "launches in the list", flush=True)
    plan = []
    n, base = positions(names, r"gemm_tc")
    # gemm-relative order inside a step: patch_embed, then 4 per trunk block (qkv, proj, fc1, fc2) ...
    plan.append(("gemm_block3", "gemm_tc", base + 1 + 4 * 3, 4))       # the first global block: qkv, proj, fc1, fc2
    plan.append(("gemm_patch_embed", "gemm_tc", base, 1))
    plan.append(("gemm_tail", "gemm_tc", base + n - 2, 2))             # un-patchify: convT_A, convT_B
    n, base = positions(names, r"attn_tc4")
    plan.append(("attn", "attn_tc4", base, 4))                         # windows 24x24, 12x48, 48x12 (padded), global
    n, base = positions(names, r"layernorm")
    plan.append(("layernorm", "layernorm", base + 6, 2))
    n, base = positions(names, r"frame_to_patches")
    plan.append(("frame_to_patches", "frame_to_patches", base, 1))
    n, base = positions(names, r"gc_quantize")
    plan.append(("quantize", "gc_quantize", base, 2))                  # encode side (symbols + indexes), decode side (indexes)
    n, base = positions(names, r"rans_encode_smem")
    plan.append(("rans_enc", "rans_encode_smem", base, 2))             # z, y
    n, base = positions(names, r"rans_decode_smem")
    plan.append(("rans_dec", "rans_decode_smem", base, 2))
    n, base = positions(names, r"attn_mma|attn_small")
    plan.append(("attn_hyper", "attn_mma|attn_small", base, 1))
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]   # e.g. --only=gemm_tail: re-capture one group
    for tag, rgx, skip, cnt in plan:
        if only and tag not in only:
            continue
        print("capture", tag, rgx, skip, cnt, flush=True)
        capture(rnd, tag, rgx, skip, cnt)
    sass_histogram(rnd)
    print("done", sorted(os.listdir(OUT))[:80], flush=True)


SITE_NAMES = {   # capture tag + position -> the name the tracked summary carries
    "gemm_block3_00": "qkv", "gemm_block3_01": "proj", "gemm_block3_02": "fc1", "gemm_block3_03": "fc2",
    "gemm_patch_embed_00": "patch_embed", "gemm_tail_00": "convT_A", "gemm_tail_01": "convT_B",
    "attn_00": "attn_window24x24", "attn_01": "attn_window12x48", "attn_02": "attn_window48x12", "attn_03": "attn_global",
    "attn_hyper_00": "attn_hyperprior", "layernorm_00": "layernorm", "frame_to_patches_00": "frame_to_patches",
    "quantize_00": "quantize_encode", "quantize_01": "quantize_decode_idx",
    "rans_enc_00": "rans_enc_z", "rans_enc_01": "rans_enc", "rans_dec_00": "rans_dec_z", "rans_dec_01": "rans_dec",
}


def collect(rnd):
    """build container: gpurun_out/prof -> profiles/ (tracked): per-site ncu summaries under their site names,
    <round>_traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic), the launch list and its summary
    over the timed step, the bench line, the SASS opcode histogram"""
    dst = os.path.join(ROOT, "profiles")
    traffic, n = collections.OrderedDict(), 0
    summaries = {}
    for f in sorted(os.listdir(OUT)):
        if f.startswith(rnd + "_") and f.endswith("_summary.json"):
            summaries.update(json.load(open(os.path.join(OUT, f))))
    for f in sorted(os.listdir(OUT)):
        m = re.match(rf"{rnd}_((?:[a-z0-9]+_)+?\d\d)_.*\.txt$", f)
        if not m or m.group(1) not in SITE_NAMES:
            continue
        site = SITE_NAMES[m.group(1)]
        shutil.copyfile(os.path.join(OUT, f), os.path.join(dst, f"{rnd}_ncu_{site}.txt"))
        n += 1
        key = f[:-4]
        if key in summaries:
            traffic[site] = dict(summaries[key], frames_per_launch=8)
    if traffic:
        json.dump(traffic, open(os.path.join(dst, f"{rnd}_traffic.json"), "w"), indent=1)
    for src, name in ((f"bench_{rnd}.json", f"{rnd}_bench.json"), (f"launches_{rnd}.csv", f"{rnd}_launches.csv"),
                      (f"{rnd}_sass_opcodes.txt", f"{rnd}_sass_opcodes.txt")):
        if os.path.exists(os.path.join(OUT, src)):
            if src.startswith("bench_"):
                line = [l for l in open(os.path.join(OUT, src)) if l.startswith("{")][-1]
                json.dump(json.loads(line), open(os.path.join(dst, name), "w"), indent=1)
            else:
                shutil.copyfile(os.path.join(OUT, src), os.path.join(dst, name))
            n += 1
    lc = os.path.join(OUT, f"launches_{rnd}.csv")
    if os.path.exists(lc):
        rows = list(csv.reader(l for l in open(lc) if l.startswith('"')))
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        per_step = (len(rows) - 1) // 4                       # 3 warm-up steps + the timed one
        agg = collections.OrderedDict()
        for row in rows[1 + 3 * per_step:]:
            name = row[ki].split("(")[0].replace("void ", "")
            v = float(row[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(row[ui], 1e-6)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        md = [f"# ncu launch list, round {rnd[1:]}", "",
              "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:<library kernels> python bench.py --steps 1 "
              "--warmup 3 --no-cpu-baseline --no-kernel-profile --no-e2e` (batch 8): the launches of the ONE timed step = 8 frames "
              f"(cold-cache, serialised under the profiler: compare SHARES with the bench's own CUDA-event table in {rnd}_bench.json, "
              "not absolutes).", "",
              f"{per_step} launches per step of 8 frames, {tot:.2f} ms summed ({tot / 8:.2f} ms per frame)", "",
              "| kernel | launches | ms | share |", "|---|---|---|---|"]
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            md.append(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f} % |")
        open(os.path.join(dst, f"{rnd}_launch_summary.md"), "w").write("\n".join(md) + "\n")
        n += 1
    print("wrote", n, "files to profiles/")


if __name__ == "__main__":
    if "--collect" in sys.argv:
        collect(sys.argv[-1])
    else:
        main(next((a for a in sys.argv[1:] if not a.startswith("--")), "r2"))
