for poly in 0 2 3 4 5; do echo "== POLY $poly"; CRA5_ATTN_POLY=$poly timeout 120 python - <<'PY'
import sys; sys.argv=['x','none']
sys.path.insert(0,'tools')
import perf_kernels as P
P.attn(16,1,10368); P.attn(16,18,576)
PY
done
ncu --set full --import-source on --clock-control none -k regex:attn_tc4 -s 1 -c 1 -o gpurun_out/attn4_global -f python tools/one_kernel.py attn 16 1 10368 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
