set -x
export OUT=${OUT:-r2h}
mkdir -p gpurun_out/${OUT:-r2h}
(timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_trees.py -q --timeout 100 -s) > gpurun_out/${OUT:-r2h}/pytest_dense.txt 2>&1
tail -15 gpurun_out/${OUT:-r2h}/pytest_dense.txt
(timeout 900 python -m pytest tests -m gpu -q --timeout 120) > gpurun_out/${OUT:-r2h}/pytest_gpu.txt 2>&1
tail -25 gpurun_out/${OUT:-r2h}/pytest_gpu.txt
(timeout 600 python bench.py --steps 20 --warmup 3) > gpurun_out/${OUT:-r2h}/bench_default.json 2> gpurun_out/${OUT:-r2h}/bench_default.err
tail -5 gpurun_out/${OUT:-r2h}/bench_default.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/'+__import__('os').environ.get('OUT','r2h')+'/bench_default.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])
    print(d['p50_step_latency_us'])
    for r in d.get('configs',[]): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('workload','batch','ms_per_launch','roofline_frac','e2e_events_per_s','e2e_ms_per_call','error')})
    for r in d['ring'].get('native',[]): print(r)
    print(d['ring'].get('run_events'), d['ring'].get('emit_await_one_caller_us'), d['ring'].get('error'))
except Exception as e: print("bench parse failed", e)
PY
(cd profiles/lab/bin && TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 suppressions=../tsan.supp" LD_LIBRARY_PATH=/usr/local/cuda/lib64:. timeout 300 setarch $(uname -m) -R ./ring_tsan 3 16) > gpurun_out/${OUT:-r2h}/tsan_ring.txt 2>&1
tail -15 gpurun_out/${OUT:-r2h}/tsan_ring.txt
grep -c "WARNING: ThreadSanitizer" gpurun_out/${OUT:-r2h}/tsan_ring.txt
(timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/lab/sanitize_smoke.py) > gpurun_out/${OUT:-r2h}/memcheck.txt 2>&1
tail -12 gpurun_out/${OUT:-r2h}/memcheck.txt
(timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python profiles/lab/sanitize_smoke.py) > gpurun_out/${OUT:-r2h}/racecheck.txt 2>&1
tail -12 gpurun_out/${OUT:-r2h}/racecheck.txt
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:dense_head -s 3 -c 1 -o gpurun_out/${OUT:-r2h}/dense python -m pytest tests/test_gpu_dense.py -q -k beats) > gpurun_out/${OUT:-r2h}/ncu_dense.log 2>&1
tail -5 gpurun_out/${OUT:-r2h}/ncu_dense.log
# A/B: top tree levels from the constant bank (default) against shared memory
for v in 1 0; do
  (B2S_T3_TOPC=$v timeout 300 python bench.py --workload trees_ens4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-configs) > gpurun_out/${OUT:-r2h}/trees_topc$v.json 2>gpurun_out/${OUT:-r2h}/trees_topc$v.err
  python -c "
import json;d=json.loads(open('gpurun_out/${OUT:-r2h}/trees_topc$v.json').read().strip().splitlines()[-1]);print('topc$v',d['ms_per_step'],d.get('kernel'),d['roofline'])" || tail -3 gpurun_out/${OUT:-r2h}/trees_topc$v.err
done
