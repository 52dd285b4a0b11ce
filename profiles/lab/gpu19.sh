# r2t (1 GPU): final single-GPU record -- trees / dense after the build-flag fix, linear with the per-NS barrier default, default bench line, launch list
set -x
export OUT=${OUT:-r2t}
mkdir -p gpurun_out/$OUT
for w in trees_ens4 dense_ens12 flow3_linear enrich_ens4; do
  (timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-configs --no-cpu-baseline) > gpurun_out/$OUT/bench_$w.json 2> gpurun_out/$OUT/bench_$w.err
  python -c "
import json;d=json.loads(open('gpurun_out/$OUT/bench_$w.json').read().strip().splitlines()[-1]);print('$w',round(d['value']/1e9,3),'G/s kernel_ms',round(d['roofline']['kernel_ms_per_launch'],5),'frac',round(d['roofline']['frac'],4),'e2e',d.get('e2e',{}).get('value'))" || tail -3 gpurun_out/$OUT/bench_$w.err
done
(timeout 900 python -m pytest tests -m gpu -q --timeout 200) > gpurun_out/$OUT/pytest_gpu.txt 2>&1
tail -4 gpurun_out/$OUT/pytest_gpu.txt
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/$OUT/smoke.txt 2>&1; tail -2 gpurun_out/$OUT/smoke.txt
(timeout 800 python bench.py) > gpurun_out/$OUT/bench_default.json 2> gpurun_out/$OUT/bench_default.err
python - <<'PY'
import json,os
try:
    d=json.loads(open('gpurun_out/'+os.environ['OUT']+'/bench_default.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['cpu_baseline']['value'])
    for r in d.get('configs',[]): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('workload','batch','ms_per_launch','roofline_frac','e2e_events_per_s')})
except Exception as e: print("bench parse failed", e)
PY
(timeout 300 python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/$OUT/bench_reference.json 2> gpurun_out/$OUT/bench_reference.err; tail -c 600 gpurun_out/$OUT/bench_reference.json
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$OUT/launches_default.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline) > gpurun_out/$OUT/launches_default.log 2>&1
bash profiles/lab/ncu_txt.sh rowthread rowthread 1048576 python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e
