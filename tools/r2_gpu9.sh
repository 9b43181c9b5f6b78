set -x
for i in 1 2 3; do
FGQ_CODEC_FUSED=0 timeout 300 python tools/host_call_timing.py 2>&1 | tail -1
FGQ_CODEC_FUSED=1 timeout 300 python tools/host_call_timing.py 2>&1 | tail -1
done
FGQ_HOST_THREADS=12 FGQ_CODEC_FUSED=1 timeout 300 python tools/host_call_timing.py 2>&1 | tail -1
FGQ_HOST_THREADS=14 FGQ_CODEC_FUSED=1 timeout 300 python tools/host_call_timing.py 2>&1 | tail -1
FGQ_CODEC_FUSED=1 timeout 900 python -m pytest tests/test_full_size_gpu.py -m gpu -q --timeout=600 -k "host_returning or host" 2>&1 | tail -3
