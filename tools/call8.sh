CRA5_GEMM_PAIR=0 timeout 120 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/gemm_single -f python tools/one_kernel.py gemm 10368 3072 1024 1 > /dev/null 2>&1
CRA5_GEMM_PAIR=1 timeout 120 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/gemm_pair -f python tools/one_kernel.py gemm 10368 3072 1024 1 > /dev/null 2>&1
ls gpurun_out/*.ncu-rep
