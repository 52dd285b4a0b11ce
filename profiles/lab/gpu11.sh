# r2l: DMMA row kernel -- parity, A/B against the DFMA kernel, warps x stages sweep, default bench line, ncu summaries
set -x
export OUT=${OUT:-r2l}
mkdir -p gpurun_out/$OUT
(timeout 600 python -m pytest tests/test_gpu_rowmma.py tests/test_gpu_parity.py tests/test_gpu_serving.py -q --timeout 200 -x) > gpurun_out/$OUT/pytest_rowmma.txt 2>&1
tail -15 gpurun_out/$OUT/pytest_rowmma.txt
one() {  # label, env...
  label=$1; shift
  (env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/ab_$label.json 2> gpurun_out/$OUT/ab_$label.err
  python -c "
import json;d=json.loads(open('gpurun_out/$OUT/ab_$label.json').read().strip().splitlines()[-1]);print('$label',round(d['value']/1e9,3),'G ev/s kernel_ms',round(d['roofline']['kernel_ms_per_launch'],5),'frac',round(d['roofline']['frac'],4),d['roofline']['kernel'][:60])" || tail -3 gpurun_out/$OUT/ab_$label.err
}
one dfma B2S_RT_MMA=0
one mma_w8s3 B2S_RM_WARPS=8 B2S_RM_STAGES=3
one mma_w6s4 B2S_RM_WARPS=6 B2S_RM_STAGES=4
one mma_w12s2 B2S_RM_WARPS=12 B2S_RM_STAGES=2
one mma_w10s2 B2S_RM_WARPS=10 B2S_RM_STAGES=2
one mma_w7s4 B2S_RM_WARPS=7 B2S_RM_STAGES=4
one mma_w4s4 B2S_RM_WARPS=4 B2S_RM_STAGES=4
(timeout 200 python bench.py --workload flow3_linear --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/ab_linear_mma.json 2>/dev/null
(B2S_RT_MMA=0 timeout 200 python bench.py --workload flow3_linear --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/ab_linear_dfma.json 2>/dev/null
python - <<'PY'
import json,os
for n in ('ab_linear_mma','ab_linear_dfma'):
    try:
        d=json.loads(open('gpurun_out/'+os.environ['OUT']+'/'+n+'.json').read().strip().splitlines()[-1]); print(n, d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel'][:50])
    except Exception as e: print(n,'failed',e)
PY
bash profiles/lab/ncu_txt.sh rowmma rowmma 1048576 python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e
date +%s > gpurun_out/$OUT/t0
(timeout 800 python bench.py) > gpurun_out/$OUT/bench_default.json 2> gpurun_out/$OUT/bench_default.err
echo "default bench wall: $(( $(date +%s) - $(cat gpurun_out/$OUT/t0) )) s"
tail -5 gpurun_out/$OUT/bench_default.err
python - <<'PY'
import json,os
try:
    d=json.loads(open('gpurun_out/'+os.environ['OUT']+'/bench_default.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline'], d['e2e'], d['cpu_baseline'])
    print(d.get('p50_step_latency_us'))
    for r in d.get('configs',[]): print({k:(round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('workload','batch','ms_per_launch','roofline_frac','e2e_events_per_s','e2e_ms_per_call','error')})
    for r in d['ring'].get('native',[]): print(r)
    print(d['ring'].get('run_events'), d['ring'].get('emit_await_one_caller_us'), d['ring'].get('error'))
except Exception as e: print("bench parse failed", e)
PY
bash profiles/lab/ncu_txt.sh dense dense_head 1048576 python bench.py --workload dense_ens12 --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e
bash profiles/lab/ncu_txt.sh gather rowthread 1048576 python bench.py --workload enrich_ens4 --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e
(timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/$OUT/launches_default.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline) > gpurun_out/$OUT/launches_default.log 2>&1
for w in dense_ens12 enrich_ens4; do
  (timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-configs) > gpurun_out/$OUT/bench_$w.json 2> gpurun_out/$OUT/bench_$w.err
done
du -sh gpurun_out; ls -la gpurun_out/$OUT
