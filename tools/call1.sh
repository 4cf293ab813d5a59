timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s2_b.json 2> gpurun_out/bench_s2_b.err; tail -3 gpurun_out/bench_s2_b.err; python - <<'PY'
import json; d=json.load(open('gpurun_out/bench_s2_b.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac']); print({k:round(v['ms_per_step'],3) for k,v in d['kernels'].items()}); print({k:(v['ms_per_step'],v['tflops']) for k,v in d['kernel_sites'].items() if v['ms_per_step']>0.1})
PY
