# r2u (1 GPU): tile / stage / grid variants of the shipped row kernel (env switches only), ingest after the copy guard
set -x
export OUT=${OUT:-r2u}
mkdir -p gpurun_out/$OUT
one() {  # label, workload, env...
  label=$1; wl=$2; shift 2
  (env "$@" timeout 200 python bench.py --workload $wl --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/ab_$label.json 2> gpurun_out/$OUT/ab_$label.err
  python -c "
import json;d=json.loads(open('gpurun_out/$OUT/ab_$label.json').read().strip().splitlines()[-1]);print('$label',round(d['value']/1e9,3),'G ev/s kernel_ms',round(d['roofline']['kernel_ms_per_launch'],5),'frac',round(d['roofline']['frac'],4),d['roofline']['kernel'][:60])" 2>/dev/null || tail -3 gpurun_out/$OUT/ab_$label.err
}
one base flow3_ens4 X=1
one tile64 flow3_ens4 B2S_RT_TILE=64
one tile64_s3 flow3_ens4 B2S_RT_TILE=64 B2S_RT_STAGES=3
one tile64_s4 flow3_ens4 B2S_RT_TILE=64 B2S_RT_STAGES=4
one gridmul2 flow3_ens4 B2S_RT_GRIDMUL=2
one gridmul4 flow3_ens4 B2S_RT_GRIDMUL=4
one tile64_gm2 flow3_ens4 B2S_RT_TILE=64 B2S_RT_GRIDMUL=2
one base_again flow3_ens4 X=1
one lin_tile64 flow3_linear B2S_RT_TILE=64
(timeout 300 python -m pytest tests/test_gpu_ingest.py -q -x --timeout 200) > gpurun_out/$OUT/pytest_ingest.txt 2>&1; tail -2 gpurun_out/$OUT/pytest_ingest.txt
(timeout 300 python bench.py --workload ingest6 --steps 10 --warmup 3 --no-configs --no-cpu-baseline) > gpurun_out/$OUT/ingest.json 2> gpurun_out/$OUT/ingest.err
python -c "
import json;d=json.loads(open('gpurun_out/$OUT/ingest.json').read().strip().splitlines()[-1]);print('ingest e2e', round(d['e2e']['value']/1e6,2),'M rows/s')"
