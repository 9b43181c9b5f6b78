set -x
timeout 300 python tools/hermite_time.py
FGQ_HERMITE_FUSED=0 timeout 300 python tools/hermite_time.py
timeout 900 python -m pytest tests/test_hermite_gpu.py tests/test_small_batch_gpu.py -m gpu -q -s --timeout=600 2>&1 | grep -i "node err\|passed\|failed\|error" | tail -12
