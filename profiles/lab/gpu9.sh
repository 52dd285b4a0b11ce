# 2-GPU validation: communicator, ShardedGraphServer, weak and strong scaling legs of bench.py
set -x
mkdir -p gpurun_out/r2i
nvidia-smi topo -m > gpurun_out/r2i/topo.txt 2>&1
(timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 300 -x) > gpurun_out/r2i/pytest_multi.txt 2>&1
tail -30 gpurun_out/r2i/pytest_multi.txt
run() {  # name, extra args...
  name=$1; shift
  (timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 "$@") > gpurun_out/r2i/$name.json 2> gpurun_out/r2i/$name.err
  tail -3 gpurun_out/r2i/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2i/$name.json').read().strip().splitlines()[-1])
    print('$name', {k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','merge','merge_verified')}, d.get('config'))
except Exception as e: print('$name parse failed', e)
PY
}
run weak2 --no-configs
run weak2_nccl --no-configs --merge nccl
run strong2_router8 --workload router8 --scaling strong --batch 65536 --no-configs
run ingest2 --workload ingest6 --no-configs
(timeout 300 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline) > gpurun_out/r2i/one.json 2> gpurun_out/r2i/one.err
(timeout 300 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline --workload router8 --batch 65536) > gpurun_out/r2i/one_router8.json 2> gpurun_out/r2i/one_router8.err
python - <<'PY'
import json
for n in ('one','one_router8'):
    try:
        d=json.loads(open(f'gpurun_out/r2i/{n}.json').read().strip().splitlines()[-1]); print(n, d['value'], d['ms_per_step'])
    except Exception as e: print(n,'failed',e)
PY
