set -x
mkdir -p gpurun_out/r2c
(timeout 600 python -m pytest tests/test_gpu_trees.py -x -q) > gpurun_out/r2c/pytest_trees.txt 2>&1
tail -30 gpurun_out/r2c/pytest_trees.txt
(timeout 600 python bench.py --workload trees_ens4 --steps 20 --warmup 3 --no-cpu-baseline) > gpurun_out/r2c/bench_trees3.json 2> gpurun_out/r2c/bench_trees3.err
tail -2 gpurun_out/r2c/bench_trees3.json | cut -c1-1800
tail -5 gpurun_out/r2c/bench_trees3.err
for w in 20 24 28; do
(B2S_T3_WARPS=$w timeout 600 python bench.py --workload trees_ens4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e) > gpurun_out/r2c/bench_trees3_w$w.json 2> gpurun_out/r2c/bench_trees3_w$w.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2c/bench_trees3_w$w.json').read().strip().splitlines()[-1]); print('W=$w', d['roofline']['kernel_ms_per_launch'], d['p50_step_latency_us']['p50'], d['config']['kernel'])"
done
(timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r2c/pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2c/pytest_gpu.txt
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:trees3 -s 4 -c 1 -o gpurun_out/r2c/trees3 python bench.py --workload trees_ens4 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e) > gpurun_out/r2c/ncu_trees3.log 2>&1
tail -3 gpurun_out/r2c/ncu_trees3.log
