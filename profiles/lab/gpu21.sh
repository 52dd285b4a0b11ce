# r2v (1 GPU, the round's last ~2 minutes of box time): the new body codec on the box's host cores (thread scaling of the pool),
# then the GPU serving tests (run_json / run_binary go through the codec) and the ingest tests
set -x
export OUT=${OUT:-r2v}
mkdir -p gpurun_out/$OUT
(timeout 50 python profiles/lab/codec_probe.py 4096 128) > gpurun_out/$OUT/codec_probe.txt 2>&1
cat gpurun_out/$OUT/codec_probe.txt
(timeout 100 python -m pytest tests/test_gpu_serving.py -q -x --timeout 80) > gpurun_out/$OUT/pytest_serving.txt 2>&1
tail -3 gpurun_out/$OUT/pytest_serving.txt
