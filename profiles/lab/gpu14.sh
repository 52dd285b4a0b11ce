# r2o: clean A/B of the fast one-hot path and the dead-tail variants (no spills), gather stages, small-batch latency with zero-copy reads
set -x
export OUT=${OUT:-r2o}
mkdir -p gpurun_out/$OUT
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rowmma.py tests/test_gpu_enrichment.py -q --timeout 200 -x) > gpurun_out/$OUT/pytest.txt 2>&1
tail -5 gpurun_out/$OUT/pytest.txt
one() {  # label, workload, env...
  label=$1; wl=$2; shift 2
  (env "$@" timeout 200 python bench.py --workload $wl --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/ab_$label.json 2> gpurun_out/$OUT/ab_$label.err
  python -c "
import json;d=json.loads(open('gpurun_out/$OUT/ab_$label.json').read().strip().splitlines()[-1]);print('$label',round(d['value']/1e9,3),'G ev/s kernel_ms',round(d['roofline']['kernel_ms_per_launch'],5),'frac',round(d['roofline']['frac'],4),d['roofline']['kernel'][:60])" 2>/dev/null || tail -3 gpurun_out/$OUT/ab_$label.err
}
one dfma_new flow3_ens4 X=1
one dfma_noskip flow3_ens4 B2S_RT_NOSKIP=1
one dfma_slowcats flow3_ens4 B2S_RT_SLOWCATS=1
one dfma_old flow3_ens4 B2S_RT_NOSKIP=1 B2S_RT_SLOWCATS=1
one dfma_new_again flow3_ens4 X=1
one dfma_new_s3 flow3_ens4 B2S_RT_STAGES=3
one dfma_new_rpt2 flow3_ens4 B2S_RT_RPT=2
one dfma_new_onesync flow3_ens4 B2S_RT_ONESYNC=1
one linear_new flow3_linear X=1
one enrich_s2 enrich_ens4 X=1
one enrich_s3 enrich_ens4 B2S_RT_STAGES=3
one enrich_s3_t64 enrich_ens4 B2S_RT_STAGES=3 B2S_RT_TILE=64
for z in 65536 1048576 4194304; do
  (B2S_ZEROCOPY_IN_BYTES=$z timeout 200 python profiles/lab/latency_probe.py) >> gpurun_out/$OUT/latency.txt 2>&1
done
cat gpurun_out/$OUT/latency.txt
bash profiles/lab/ncu_txt.sh rowthread rowthread 1048576 python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e
