set -x
mkdir -p gpurun_out/r2f
(timeout 600 python -m pytest tests/test_gpu_trees.py -x -q) > gpurun_out/r2f/pytest_trees.txt 2>&1
tail -5 gpurun_out/r2f/pytest_trees.txt
run() {
(timeout 600 python bench.py --workload trees_ens4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e) > gpurun_out/r2f/bench_$1.json 2> gpurun_out/r2f/bench_$1.err
python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/r2f/bench_$1.json').read().strip().splitlines()[-1]); print('$1', d['roofline']['kernel_ms_per_launch'], d['p50_step_latency_us']['p50'], d['config']['kernel'][-60:])
except Exception as e: print('$1 FAILED', e)"
}
run auto
B2S_T3_UNROLL=4 run auto_u4
for w in 10 13 17 20; do B2S_T3_WARPS=$w run w$w; B2S_T3_WARPS=$w B2S_T3_UNROLL=4 run w${w}_u4; done
(timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2f/pytest_gpu.txt 2>&1
tail -8 gpurun_out/r2f/pytest_gpu.txt
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:"trees3" -s 2 -c 1 -o gpurun_out/r2f/trees3 python bench.py --workload trees_ens4 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e) > gpurun_out/r2f/ncu_trees3.log 2>&1
tail -3 gpurun_out/r2f/ncu_trees3.log
