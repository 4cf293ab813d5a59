tools/micro/tanh_err > gpurun_out/tanh_err.txt 2>&1; cat gpurun_out/tanh_err.txt
for ex in 1 0; do
  CRA5_GELU_EXACT=$ex timeout 400 python bench.py --steps 6 --no-cpu-baseline --no-e2e > gpurun_out/bench_gelu_exact$ex.json 2> gpurun_out/bench_gelu_exact$ex.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_gelu_exact$ex.json").read().strip().splitlines()[-1])
print("GELU_EXACT=$ex value", round(d["value"],2), "fc1", d["kernel_sites"]["gemm_tc:fc1"], "clk", d["clocks"].get("sm_mhz"))
PY
done
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_precision.py tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -4
