# refresh after the epilogue / tile-shape changes: launch list + one trunk block pair of GEMM captures
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-profile"
K="regex:attn_|gemm_tc|layernorm|rans_|frame_to|gc_quant|eb_quant|scan_len|compact_|container_|im2col|transpose_cast|cast_bf16|word_to"
mkdir -p /tmp/cap gpurun_out/prof
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 920 -c 320 --csv --log-file gpurun_out/prof/launches_r1.csv $B > /dev/null 2>&1
F="--set full --import-source on --clock-control none"
timeout 400 ncu $F -k regex:gemm_tc -s 649 -c 8 -o /tmp/cap/gemm -f $B > /dev/null 2>&1
rm -f gpurun_out/prof/gemm_*
python tools/ncu_summarize.py /tmp/cap/gemm.ncu-rep gpurun_out/prof gemm
wc -l gpurun_out/prof/launches_r1.csv
