set -x
mkdir -p gpurun_out/r2d
(timeout 600 python -m pytest tests/test_gpu_trees.py -x -q) > gpurun_out/r2d/pytest_trees.txt 2>&1
tail -5 gpurun_out/r2d/pytest_trees.txt
for cfg in "4 0" "7 0" "7 20" "10 0" "10 20" "12 20"; do
set -- $cfg
(B2S_T3_PROD=$1 $( [ "$2" != "0" ] && echo B2S_T3_WARPS=$2 ) timeout 600 python bench.py --workload trees_ens4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e) > gpurun_out/r2d/bench_trees3_p$1_w$2.json 2> gpurun_out/r2d/bench_trees3_p$1_w$2.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2d/bench_trees3_p$1_w$2.json').read().strip().splitlines()[-1]); print('prod=$1 W=$2', d['roofline']['kernel_ms_per_launch'], d['p50_step_latency_us']['p50'], d['config']['kernel'])"
done
(timeout 900 python bench.py --steps 20 --warmup 3) > gpurun_out/r2d/bench_default.json 2> gpurun_out/r2d/bench_default.err
tail -c 6000 gpurun_out/r2d/bench_default.json
tail -5 gpurun_out/r2d/bench_default.err
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:trees3 -s 4 -c 1 -o gpurun_out/r2d/trees3 python bench.py --workload trees_ens4 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e) > gpurun_out/r2d/ncu_trees3.log 2>&1
tail -3 gpurun_out/r2d/ncu_trees3.log
