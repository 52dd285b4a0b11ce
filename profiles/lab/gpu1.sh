set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(timeout 600 profiles/lab/bin/trees_lab 20 262144) > gpurun_out/r2a/trees_lab.txt 2>&1
tail -40 gpurun_out/r2a/trees_lab.txt
(timeout 600 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2a/trees_lab_ncu.csv profiles/lab/bin/trees_lab 1 262144) > gpurun_out/r2a/trees_lab_ncu.log 2>&1
(timeout 900 python bench.py --workload trees_ens4 --steps 20 --warmup 3) > gpurun_out/r2a/bench_trees.json 2> gpurun_out/r2a/bench_trees.err
tail -3 gpurun_out/r2a/bench_trees.json
(timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r2a/pytest_gpu.txt 2>&1
tail -5 gpurun_out/r2a/pytest_gpu.txt
