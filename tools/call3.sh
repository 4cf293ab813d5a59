for cfg in "1 3 0" "1 3 1" "2 3 1"; do set -- $cfg; echo "== SPLIT $1 POLY $2 SPIN $3"; CRA5_ATTN_SPLIT=$1 CRA5_ATTN_POLY=$2 CRA5_ATTN_SPIN=$3 timeout 40 python - <<'PY'
import sys; sys.argv=['x','none']
sys.path.insert(0,'tools')
import perf_kernels as P
P.attn(16,1,10368); P.attn(16,18,576); P.attn(16,24,576)
PY
done
