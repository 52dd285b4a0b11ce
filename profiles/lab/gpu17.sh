# r2r (2 GPUs): after the fixes (upload ordered on the library stream, one fence per CTA, stream-memop wait)
set -x
export OUT=${OUT:-r2r}
NP=2
mkdir -p gpurun_out/$OUT
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 profiles/lab/multi_repro.py) > gpurun_out/$OUT/repro.txt 2>&1
grep -E "^rank" gpurun_out/$OUT/repro.txt | sort | head
(timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 400 -x) > gpurun_out/$OUT/pytest_multi.txt 2>&1
tail -5 gpurun_out/$OUT/pytest_multi.txt
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  (env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NP --steps 20 --warmup 3 --no-configs --no-cpu-baseline "$@") > gpurun_out/$OUT/$name.json 2> gpurun_out/$OUT/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$OUT/$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value']/1e9,3), 'G ev/s  ms/launch', round(d['ms_per_step']/d['config']['launches_per_step'],5), 'merge_verified', d['config'].get('merge_verified'), 'e2e', d.get('e2e',{}).get('value'))
except Exception as e: print('$name parse failed', e); print(open('gpurun_out/$OUT/$name.err').read()[-600:])
PY
}
run weak2 X=1 --
run weak2_lag0 X=1 -- --merge-lag 0
run weak2_kernelwait B2S_COMM_WAIT=kernel --
run weak2_nccl X=1 -- --merge nccl --no-e2e
run strong2_router8 X=1 -- --workload router8 --scaling strong --batch 65536
(timeout 300 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/one.json 2> gpurun_out/$OUT/one.err
(timeout 300 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e --workload router8 --batch 65536) > gpurun_out/$OUT/one_router8.json 2> gpurun_out/$OUT/one_router8.err
python - <<'PY'
import json,os
for n in ('one','one_router8'):
    try:
        d=json.loads(open('gpurun_out/'+os.environ['OUT']+'/'+n+'.json').read().strip().splitlines()[-1]); print(n, round(d['value']/1e9,3), d['ms_per_step']/d['config']['launches_per_step'])
    except Exception as e: print(n,'failed',e)
PY
