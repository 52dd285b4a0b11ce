# Re-entry validation (r2k): whole GPU suite, the driver's default bench line, dense head v2 + enrichment ncu captures
set -x
export OUT=${OUT:-r2k}
mkdir -p gpurun_out/$OUT
(timeout 900 python -m pytest tests -m gpu -q --timeout 150) > gpurun_out/$OUT/pytest_gpu.txt 2>&1
tail -25 gpurun_out/$OUT/pytest_gpu.txt
(timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > gpurun_out/$OUT/smoke.txt 2>&1; tail -3 gpurun_out/$OUT/smoke.txt
/usr/bin/time -v -o gpurun_out/$OUT/bench_default.time timeout 800 python bench.py > gpurun_out/$OUT/bench_default.json 2> gpurun_out/$OUT/bench_default.err
tail -5 gpurun_out/$OUT/bench_default.err; grep -E "Elapsed|Maximum resident" gpurun_out/$OUT/bench_default.time
python - <<'PY'
import json,os
try:
    d=json.loads(open('gpurun_out/'+os.environ['OUT']+'/bench_default.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline'], d['e2e'], d['cpu_baseline'])
    print(d.get('p50_step_latency_us'))
    for r in d.get('configs',[]): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('workload','batch','kernel','ms_per_launch','roofline_frac','e2e_events_per_s','e2e_ms_per_call','error')})
    for r in d['ring'].get('native',[]): print(r)
    print(d['ring'].get('run_events'), d['ring'].get('emit_await_one_caller_us'), d['ring'].get('error'))
except Exception as e: print("bench parse failed", e)
PY
for w in dense_ens12 enrich_ens4; do
  (timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-configs) > gpurun_out/$OUT/bench_$w.json 2> gpurun_out/$OUT/bench_$w.err
  python -c "
import json;d=json.loads(open('gpurun_out/$OUT/bench_$w.json').read().strip().splitlines()[-1]);print('$w',d['value'],d['ms_per_step'],d.get('kernel'),d['roofline'],d['e2e'])" || tail -3 gpurun_out/$OUT/bench_$w.err
done
(B2S_DENSE=0 timeout 300 python bench.py --workload dense_ens12 --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/bench_dense_off.json 2> gpurun_out/$OUT/bench_dense_off.err
python -c "
import json;d=json.loads(open('gpurun_out/$OUT/bench_dense_off.json').read().strip().splitlines()[-1]);print('dense_off',d['value'],d['ms_per_step'],d.get('kernel'),d['roofline'])"
# ncu: launch list of the default command, then full captures of the dense head, the metric kernel and the gather loader
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$OUT/launches_default.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline) > gpurun_out/$OUT/launches_default.log 2>&1
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_head -s 3 -c 1 -o gpurun_out/$OUT/dense python bench.py --workload dense_ens12 --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/ncu_dense.log 2>&1; tail -2 gpurun_out/$OUT/ncu_dense.log
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowthread -s 3 -c 1 -o gpurun_out/$OUT/rowthread python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/ncu_rowthread.log 2>&1; tail -2 gpurun_out/$OUT/ncu_rowthread.log
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:rowthread -s 3 -c 1 -o gpurun_out/$OUT/gather python bench.py --workload enrich_ens4 --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/ncu_gather.log 2>&1; tail -2 gpurun_out/$OUT/ncu_gather.log
ls -la gpurun_out/$OUT
