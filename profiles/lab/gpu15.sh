# r2p (2 GPUs): bisect the sharded-server mismatch; what a merged step costs (lab switches); ingest with 2-D column copies
set -x
export OUT=${OUT:-r2p}
NP=2
mkdir -p gpurun_out/$OUT
rep() {  # label, env...
  label=$1; shift
  (env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 profiles/lab/multi_repro.py) > gpurun_out/$OUT/repro_$label.txt 2>&1
  echo "== $label"; grep -E "^rank" gpurun_out/$OUT/repro_$label.txt | sort | head -24
}
rep default X=1
rep onesync0 B2S_RT_ONESYNC=0
rep noskip B2S_RT_NOSKIP=1
rep slowcats B2S_RT_SLOWCATS=1
run() {  # name, env..., -- args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  (env "${envs[@]}" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NP --steps 10 --warmup 3 --no-configs --no-cpu-baseline --no-e2e "$@") > gpurun_out/$OUT/$name.json 2> gpurun_out/$OUT/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$OUT/$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value']/1e9,3), 'G ev/s  ms/launch', round(d['ms_per_step']/d['config']['launches_per_step'],5), 'merge_verified', d['config'].get('merge_verified'))
except Exception as e: print('$name parse failed', e); print(open('gpurun_out/$OUT/$name.err').read()[-600:])
PY
}
run weak2 X=1 --
run weak2_nowait B2S_LAB_COMM_NOWAIT=1 --
run weak2_nosignal B2S_LAB_COMM_NOWAIT=1 B2S_LAB_COMM_NOSIGNAL=1 --
run weak2_selfonly B2S_LAB_COMM_NOWAIT=1 B2S_LAB_COMM_NOSIGNAL=1 B2S_LAB_COMM_SELFONLY=1 --
run weak2_selfonly_sig B2S_LAB_COMM_SELFONLY=1 --
