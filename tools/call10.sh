timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for poly in 2 3 4; do echo "== POLY $poly"; CRA5_ATTN_POLY=$poly timeout 60 python - <<'PY'
import sys; sys.argv=['x','none']
sys.path.insert(0,'tools')
import perf_kernels as P
P.attn(16,1,10368); P.attn(16,18,576); P.attn(16,24,576)
PY
done
