for s in 1 2; do echo "== SPLIT $s"; CRA5_ATTN_SPLIT=$s timeout 120 python tools/determinism.py model 2>&1 | grep -v serving; done
echo "== v3"; CRA5_ATTN=3 timeout 120 python tools/determinism.py model 2>&1 | grep -v serving
