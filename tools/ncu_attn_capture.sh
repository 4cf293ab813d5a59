#!/bin/bash
# One --set full --import-source on capture of the global-shape attention launch (tools/perf_attn.py's child process),
# kept as a .ncu-rep under gpurun_out/ so that the per-instruction stall samples can be read in the build container:
#   ncu -i gpurun_out/attn_global.ncu-rep --page source --csv > /tmp/attn_source.csv
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attn_tc4 -s 7 -c 1 -o gpurun_out/attn_global -f \
    python tools/perf_attn.py --child > gpurun_out/ncu_attn_capture.log 2>&1
ls -la gpurun_out/attn_global.ncu-rep
