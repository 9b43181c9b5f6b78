set -x
timeout 300 python tools/hermite_time.py
timeout 300 python tools/laguerre_time.py > gpurun_out/r2_lag_airy.txt 2> gpurun_out/r2_lag_airy_t.txt; cat gpurun_out/r2_lag_airy_t.txt
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_airy.txt 2>&1
timeout 900 python -m pytest tests/test_hermite_gpu.py tests/test_laguerre_gpu.py -m gpu -q --timeout=600 2>&1 | tail -3
