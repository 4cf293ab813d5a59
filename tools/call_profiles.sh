# round-1 evidence: launch list of the bench command, then --set full captures of the kernels that matter.
# Reports stay in /tmp on the GPU box; only text summaries come back under gpurun_out/prof/.
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-profile"
mkdir -p /tmp/cap gpurun_out/prof
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cra5 -s 900 -c 320 --csv --log-file gpurun_out/prof/launches_r1.csv $B > /dev/null 2>&1
F="--set full --import-source on --clock-control none"
timeout 400 ncu $F -k regex:gemm_tc_kernel -s 651 -c 8 -o /tmp/cap/gemm -f $B > /dev/null 2>&1
python tools/ncu_summarize.py /tmp/cap/gemm.ncu-rep gpurun_out/prof gemm
timeout 400 ncu $F -k "regex:attn_tc4|layernorm_bf16|frame_to_patches" -s 312 -c 9 -o /tmp/cap/attn -f $B > /dev/null 2>&1
python tools/ncu_summarize.py /tmp/cap/attn.ncu-rep gpurun_out/prof attn
timeout 400 ncu $F -k "regex:rans_.*smem|gc_quantize|attn_mma" -s 54 -c 8 -o /tmp/cap/entropy -f $B > /dev/null 2>&1
python tools/ncu_summarize.py /tmp/cap/entropy.ncu-rep gpurun_out/prof entropy
du -sh gpurun_out/prof
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/prof/bench_r1.json 2> gpurun_out/prof/bench_r1.err; tail -2 gpurun_out/prof/bench_r1.err
