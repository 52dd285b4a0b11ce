# r2n: 2-GPU validation: communicator, ShardedGraphServer, weak and strong scaling legs of bench.py
set -x
export OUT=${OUT:-r2n}
NP=${NP:-2}
mkdir -p gpurun_out/$OUT
nvidia-smi topo -m > gpurun_out/$OUT/topo.txt 2>&1
(timeout 600 python -m pytest tests/test_gpu_multi.py -q --timeout 300 -x) > gpurun_out/$OUT/pytest_multi.txt 2>&1
tail -30 gpurun_out/$OUT/pytest_multi.txt
run() {  # name, extra args...
  name=$1; shift
  (timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NP --steps 20 --warmup 3 "$@") > gpurun_out/$OUT/$name.json 2> gpurun_out/$OUT/$name.err
  tail -3 gpurun_out/$OUT/$name.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/$OUT/$name.json').read().strip().splitlines()[-1])
    print('$name', {k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','merge','merge_verified')}, d.get('config'))
except Exception as e: print('$name parse failed', e)
PY
}
run weak2 --no-configs
run weak2_lag0 --no-configs --merge-lag 0
run weak2_nccl --no-configs --merge nccl
run strong2_router8 --workload router8 --scaling strong --batch 65536 --no-configs
run strong2_router8_lag0 --workload router8 --scaling strong --batch 65536 --no-configs --merge-lag 0
run ingest2 --workload ingest6 --no-configs
(timeout 300 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline) > gpurun_out/$OUT/one.json 2> gpurun_out/$OUT/one.err
(timeout 300 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu-baseline --workload router8 --batch 65536) > gpurun_out/$OUT/one_router8.json 2> gpurun_out/$OUT/one_router8.err
python - <<'PY'
import json
for n in ('one','one_router8'):
    try:
        d=json.loads(open('gpurun_out/'+__import__('os').environ['OUT']+'/'+n+'.json').read().strip().splitlines()[-1]); print(n, d['value'], d['ms_per_step'])
    except Exception as e: print(n,'failed',e)
PY

# NVLink counters of the scoring kernel with the fused merge stores (rank 0 under ncu, the peers only hold their buffers)
(MASTER_ADDR=127.0.0.1 MASTER_PORT=29533 WORLD_SIZE=2 RANK=1 timeout 200 python profiles/lab/nvlink_probe.py > gpurun_out/$OUT/nvlink_rank1.log 2>&1 &)
(MASTER_ADDR=127.0.0.1 MASTER_PORT=29533 WORLD_SIZE=2 RANK=0 timeout 200 ncu --clock-control none -k regex:rowthread -s 2 -c 3 --metrics nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes.sum.per_second,lts__t_sectors_srcunit_tex_aperture_peer_op_write.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/$OUT/nvlink_ncu.csv python profiles/lab/nvlink_probe.py) > gpurun_out/$OUT/nvlink_rank0.log 2>&1
tail -3 gpurun_out/$OUT/nvlink_rank0.log; tail -12 gpurun_out/$OUT/nvlink_ncu.csv | cut -c1-300
