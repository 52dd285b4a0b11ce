# r2w (1 GPU, what is left of the round's box time): the full GPU suite on the final code, the wire path with the new body
# parser (trees_ens4 line carries "wire"), and the default bench line
set -x
export OUT=${OUT:-r2w}
mkdir -p gpurun_out/$OUT
(timeout 110 python -m pytest tests -m gpu -q -x --timeout 90) > gpurun_out/$OUT/pytest_gpu.txt 2>&1
tail -3 gpurun_out/$OUT/pytest_gpu.txt
(timeout 45 python bench.py --workload trees_ens4 --steps 10 --warmup 3 --no-configs --no-cpu-baseline) > gpurun_out/$OUT/bench_trees_ens4.json 2> gpurun_out/$OUT/bench_trees_ens4.err
python -c "
import json;d=json.loads(open('gpurun_out/$OUT/bench_trees_ens4.json').read().strip().splitlines()[-1]);print('wire', d.get('wire'))" || tail -3 gpurun_out/$OUT/bench_trees_ens4.err
(timeout 75 python bench.py) > gpurun_out/$OUT/bench_default.json 2> gpurun_out/$OUT/bench_default.err
python -c "
import json;d=json.loads(open('gpurun_out/$OUT/bench_default.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'])" || tail -3 gpurun_out/$OUT/bench_default.err
