# r2m: DFMA row kernel with the fast one-hot path and chunk skipping (A/B), gather-loader stage/tile sweep, ncu of the default kernel
set -x
export OUT=${OUT:-r2m}
mkdir -p gpurun_out/$OUT
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rowmma.py tests/test_gpu_enrichment.py tests/test_gpu_serving.py -q --timeout 200 -x) > gpurun_out/$OUT/pytest.txt 2>&1
tail -15 gpurun_out/$OUT/pytest.txt
one() {  # label, workload, env...
  label=$1; wl=$2; shift 2
  (env "$@" timeout 200 python bench.py --workload $wl --steps 20 --warmup 3 --no-configs --no-cpu-baseline --no-e2e) > gpurun_out/$OUT/ab_$label.json 2> gpurun_out/$OUT/ab_$label.err
  python -c "
import json;d=json.loads(open('gpurun_out/$OUT/ab_$label.json').read().strip().splitlines()[-1]);print('$label',round(d['value']/1e9,3),'G ev/s kernel_ms',round(d['roofline']['kernel_ms_per_launch'],5),'frac',round(d['roofline']['frac'],4),d['roofline']['kernel'][:60])" || tail -3 gpurun_out/$OUT/ab_$label.err
}
one dfma_new flow3_ens4 X=1
one dfma_noskip flow3_ens4 B2S_RT_NOSKIP=1
one dfma_slowcats flow3_ens4 B2S_RT_SLOWCATS=1
one dfma_old flow3_ens4 B2S_RT_NOSKIP=1 B2S_RT_SLOWCATS=1
one dfma_new_tpr1 flow3_ens4 B2S_RT_TPR=1
one dfma_new_s3 flow3_ens4 B2S_RT_STAGES=3
one dfma_new_rpt2 flow3_ens4 B2S_RT_RPT=2
one mma_w12s2_new flow3_ens4 B2S_RT_MMA=1
one linear_new flow3_linear X=1
for st in 2 3 4; do for tile in 128 64; do
  one enrich_s${st}_t${tile} enrich_ens4 B2S_RT_STAGES=$st B2S_RT_TILE=$tile
done; done
bash profiles/lab/ncu_txt.sh rowthread rowthread 1048576 python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-e2e
du -sh gpurun_out; ls gpurun_out/$OUT | wc -l
