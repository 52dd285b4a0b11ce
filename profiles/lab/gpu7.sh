set -x
mkdir -p gpurun_out/r2g
(timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2g/pytest_gpu.txt 2>&1
tail -12 gpurun_out/r2g/pytest_gpu.txt
(timeout 900 python bench.py --steps 20 --warmup 3) > gpurun_out/r2g/bench_default.json 2> gpurun_out/r2g/bench_default.err
tail -5 gpurun_out/r2g/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2g/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
print(d['p50_step_latency_us'])
for r in d.get('configs',[]): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('workload','batch','ms_per_launch','roofline_frac','e2e_events_per_s','e2e_ms_per_call','error')})
for r in d['ring'].get('native',[]): print(r)
print(d['ring'].get('run_events'), d['ring'].get('emit_await_one_caller_us'), d['ring'].get('error'))
PY
