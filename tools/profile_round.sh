#!/bin/bash
# One-call profiling recipe for a round (run ON the GPU box: `gpurun --timeout 1700 -- bash tools/profile_round.sh r2`).
#   1. pytest -m gpu (parity first), 2. the bench line, 3. the ncu launch list of the same bench command
#   (gpu__time_duration.sum, --clock-control none: cold-cache, serialised -- compare SHARES with the bench's own
#   CUDA-event table, not absolutes), 4. ncu --set full captures of one trunk block's GEMMs, the attention kernels and
#   the entropy kernels, summarised on the box (reports are tens of MB; only the text summaries travel back).
# Results: gpurun_out/prof/{bench_<round>.json, launches_<round>.csv, <site>_*.txt, *_summary.json};
# copy what should be judged into profiles/ with tools/summarize_profiles.py <round>.
R=${1:-r2}
EXTRA=${2:-}            # e.g. "--lanes 2"; CRA5_PDL=1 is taken from the environment
mkdir -p /tmp/cap gpurun_out/prof
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 500 python bench.py --steps 10 --warmup 3 ${EXTRA} > gpurun_out/prof/bench_${R}.json 2> gpurun_out/prof/bench_${R}.err
tail -2 gpurun_out/prof/bench_${R}.err
python - <<PY
import json
d = json.load(open("gpurun_out/prof/bench_${R}.json"))
print("value", d["value"], "e2e", d["e2e"]["value"] if d.get("e2e") else None, "roofline", d["roofline"]["frac"] if d.get("roofline") else None)
print({k: round(v["ms_per_step"], 3) for k, v in (d.get("kernels") or {}).items()})
PY
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-profile ${EXTRA}"
K="regex:attn_|gemm_tc|layernorm|rans_|frame_to|gc_quant|eb_quant|scan_len|compact_|container_|im2col|transpose_cast|cast_bf16|word_to"
# launch list: skip the warm-up frames (about 300 launches per frame), keep one timed frame
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 920 -c 320 --csv \
    --log-file gpurun_out/prof/launches_${R}.csv $B > /dev/null 2>&1
F="--set full --import-source on --clock-control none"
timeout 400 ncu $F -k regex:gemm_tc -s 649 -c 8 -o /tmp/cap/gemm -f $B > /dev/null 2>&1
python tools/ncu_summarize.py /tmp/cap/gemm.ncu-rep gpurun_out/prof gemm
timeout 400 ncu $F -k regex:attn_tc4 -s 75 -c 4 -o /tmp/cap/attn -f $B > /dev/null 2>&1
python tools/ncu_summarize.py /tmp/cap/attn.ncu-rep gpurun_out/prof attn
timeout 400 ncu $F -k "regex:rans_.*smem|gc_quantize|layernorm_bf16_vec" -s 18 -c 8 -o /tmp/cap/entropy -f $B > /dev/null 2>&1
python tools/ncu_summarize.py /tmp/cap/entropy.ncu-rep gpurun_out/prof entropy
wc -l gpurun_out/prof/launches_${R}.csv
