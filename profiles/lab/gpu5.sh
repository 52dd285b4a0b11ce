set -x
mkdir -p gpurun_out/r2e
(timeout 600 python -m pytest tests/test_gpu_trees.py -x -q) > gpurun_out/r2e/pytest_trees.txt 2>&1
tail -15 gpurun_out/r2e/pytest_trees.txt
for w in 0 13 17 20 28; do
if [ "$w" = "0" ]; then unset B2S_T3_WARPS; else export B2S_T3_WARPS=$w; fi
(timeout 600 python bench.py --workload trees_ens4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e) > gpurun_out/r2e/bench_trees3_w$w.json 2> gpurun_out/r2e/bench_trees3_w$w.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2e/bench_trees3_w$w.json').read().strip().splitlines()[-1]); print('W=$w', d['roofline']['kernel_ms_per_launch'], d['p50_step_latency_us']['p50'], d['config']['kernel'])"
done
unset B2S_T3_WARPS
(timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/r2e/pytest_gpu.txt 2>&1
tail -8 gpurun_out/r2e/pytest_gpu.txt
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:"trees3|t3_prep" -s 8 -c 2 -o gpurun_out/r2e/trees3 python bench.py --workload trees_ens4 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e) > gpurun_out/r2e/ncu_trees3.log 2>&1
tail -3 gpurun_out/r2e/ncu_trees3.log
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/r2e/launches_trees.csv python bench.py --workload trees_ens4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e) > /dev/null 2>&1
tail -12 gpurun_out/r2e/launches_trees.csv | cut -c1-200
