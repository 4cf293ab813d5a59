#!/bin/bash
# One GPU-box call for a round's refresh: the whole -m gpu suite, smoke, the default bench line, the batch-1 timeline.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gputest_r2_final.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputest_r2_final.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_r2_final.txt 2>&1; tail -1 gpurun_out/smoke_r2_final.txt
timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r2_final.json
timeout 300 python tools/step_timeline.py --batch 1 > gpurun_out/step_timeline_b1.txt 2>&1; tail -6 gpurun_out/step_timeline_b1.txt
