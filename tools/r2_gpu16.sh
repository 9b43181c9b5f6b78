set -x
timeout 300 python tools/small_batch_timing.py 2>&1 | tail -4
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_pack.txt 2>&1
timeout 900 python -m pytest tests/test_small_batch_gpu.py tests/test_laguerre_gpu.py tests/test_hermite_gpu.py tests/test_jacobi_gpu.py -m gpu -q --timeout=600 2>&1 | tail -4
timeout 300 python tools/laguerre_time.py 2>&1 >/dev/null | head -14
