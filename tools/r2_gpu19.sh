set -x
timeout 300 python tools/hermite_time.py
timeout 300 python tools/laguerre_time.py 2>&1 >/dev/null | sed -n '9,12p'
timeout 900 python -m pytest tests/test_hermite_gpu.py tests/test_laguerre_gpu.py tests/test_small_batch_gpu.py -m gpu -q --timeout=600 2>&1 | tail -3
