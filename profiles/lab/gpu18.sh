# r2s (2 GPUs): fused wait (the launch's last CTA waits for the flags) against the wait kernel; the multi-GPU tests
set -x
export OUT=${OUT:-r2s}
NP=2
mkdir -p gpurun_out/$OUT
(timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 500 -x) > gpurun_out/$OUT/pytest_multi.txt 2>&1
tail -5 gpurun_out/$OUT/pytest_multi.txt; grep -n "rank0\]:" gpurun_out/$OUT/pytest_multi.txt | head -8
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  (env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NP --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e "$@") > gpurun_out/$OUT/$name.json 2> gpurun_out/$OUT/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$OUT/$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value']/1e9,3), 'G ev/s  ms/launch', round(d['ms_per_step']/d['config']['launches_per_step'],5), 'merge_verified', d['config'].get('merge_verified'))
except Exception as e: print('$name parse failed', e); print(open('gpurun_out/$OUT/$name.err').read()[-600:])
PY
}
run weak2_fused_lag1 X=1 --
run weak2_fused_lag0 X=1 -- --merge-lag 0
run weak2_kernel_lag1 X=1 -- --merge-wait kernel
run strong2_router8_fused X=1 -- --workload router8 --scaling strong --batch 65536
run strong2_router8_kernel X=1 -- --workload router8 --scaling strong --batch 65536 --merge-wait kernel
(timeout 300 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/one.json 2> gpurun_out/$OUT/one.err
python - <<'PY'
import json,os
for n in ('one',):
    try:
        d=json.loads(open('gpurun_out/'+os.environ['OUT']+'/'+n+'.json').read().strip().splitlines()[-1]); print(n, round(d['value']/1e9,3), d['ms_per_step']/d['config']['launches_per_step'])
    except Exception as e: print(n,'failed',e)
PY
