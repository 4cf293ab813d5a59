timeout 300 ncu --set full --import-source on --clock-control none -k regex:rans_.*smem -s 4 -c 4 -o gpurun_out/rans_smem -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-kernel-profile > /dev/null 2>&1
ls -la gpurun_out/rans_smem.ncu-rep
