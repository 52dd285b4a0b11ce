set -x
mkdir -p gpurun_out/r2b
(timeout 900 python -m pytest tests/test_gpu_trees.py -x -q) > gpurun_out/r2b/pytest_trees.txt 2>&1
tail -30 gpurun_out/r2b/pytest_trees.txt
(timeout 600 python bench.py --workload trees_ens4 --steps 20 --warmup 3 --cpu-seconds 4) > gpurun_out/r2b/bench_trees3.json 2> gpurun_out/r2b/bench_trees3.err
tail -2 gpurun_out/r2b/bench_trees3.json | cut -c1-1500
tail -5 gpurun_out/r2b/bench_trees3.err
(B2S_TREES=2 timeout 600 python bench.py --workload trees_ens4 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e) > gpurun_out/r2b/bench_trees2.json 2> gpurun_out/r2b/bench_trees2.err
tail -2 gpurun_out/r2b/bench_trees2.json | cut -c1-600
(timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/r2b/pytest_gpu.txt 2>&1
tail -15 gpurun_out/r2b/pytest_gpu.txt
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:trees3 -s 4 -c 1 -o gpurun_out/r2b/trees3 python bench.py --workload trees_ens4 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e) > gpurun_out/r2b/ncu_trees3.log 2>&1
tail -3 gpurun_out/r2b/ncu_trees3.log
ls -la gpurun_out/r2b
