# r2q (1 GPU): ingest with 2-D column copies (A/B), whole GPU suite, the default bench line
set -x
export OUT=${OUT:-r2q}
mkdir -p gpurun_out/$OUT
for ch in 65536 262144; do
for d2 in 1 0; do
  (B2S_COLS_2D=$d2 B2S_COLS_CHUNK=$ch timeout 300 python bench.py --workload ingest6 --steps 10 --warmup 3 --no-configs --no-cpu-baseline) > gpurun_out/$OUT/ingest_2d${d2}_c$ch.json 2> gpurun_out/$OUT/ingest_2d${d2}_c$ch.err
  python -c "
import json;d=json.loads(open('gpurun_out/$OUT/ingest_2d${d2}_c$ch.json').read().strip().splitlines()[-1]);print('ingest 2d=$d2 chunk=$ch e2e', round(d['e2e']['value']/1e6,2),'M rows/s; dataframe', round(d['e2e']['dataframe_boundary']['value']/1e6,2))" || tail -3 gpurun_out/$OUT/ingest_2d${d2}_c$ch.err
done; done
(timeout 300 python -m pytest tests/test_gpu_ingest.py -q -x --timeout 200) > gpurun_out/$OUT/pytest_ingest.txt 2>&1; tail -3 gpurun_out/$OUT/pytest_ingest.txt
(timeout 900 python -m pytest tests -m gpu -q --timeout 200) > gpurun_out/$OUT/pytest_gpu.txt 2>&1
tail -6 gpurun_out/$OUT/pytest_gpu.txt
(timeout 800 python bench.py) > gpurun_out/$OUT/bench_default.json 2> gpurun_out/$OUT/bench_default.err
tail -3 gpurun_out/$OUT/bench_default.err
python - <<'PY'
import json,os
try:
    d=json.loads(open('gpurun_out/'+os.environ['OUT']+'/bench_default.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['cpu_baseline']['value'])
    print(d.get('p50_step_latency_us'))
    for r in d.get('configs',[]): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('workload','batch','ms_per_launch','roofline_frac','e2e_events_per_s','e2e_ms_per_call','error')})
    for r in d['ring'].get('native',[]): print(r)
    print(d['ring'].get('run_events'), d['ring'].get('emit_await_one_caller_us'), d['ring'].get('error'))
except Exception as e: print("bench parse failed", e)
PY
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$OUT/launches_default.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline) > gpurun_out/$OUT/launches_default.log 2>&1
(timeout 300 python bench.py --workload trees_ens4 --steps 20 --warmup 3 --no-configs) > gpurun_out/$OUT/bench_trees_ens4.json 2> gpurun_out/$OUT/bench_trees_ens4.err
(timeout 300 python bench.py --workload flow3_linear --steps 20 --warmup 3 --no-configs) > gpurun_out/$OUT/bench_flow3_linear.json 2> gpurun_out/$OUT/bench_flow3_linear.err
