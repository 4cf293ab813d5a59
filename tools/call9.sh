for env in "X=0" "CRA5_GEMM_BN=128" "CRA5_GEMM_BN=256" "CRA5_GEMM_PAIR=1" "CRA5_GEMM_PAIR=0"; do echo "== $env"; env $env timeout 60 python - <<'PY'
import sys; sys.argv=['x','none']
sys.path.insert(0,'tools')
import perf_kernels as P
import builtins
_p=builtins.print
def q(*a,**k):
    if a and 'cublas' in str(a[0]): return
    _p(*a,**k)
builtins.print=q
P.gemm(10368,1024,1024,4); P.gemm(13824,1024,1024,4); P.gemm(10368,1024,4096,4); P.gemm(10368,4096,1024,2); P.gemm(10368,3072,1024,1)
PY
done
