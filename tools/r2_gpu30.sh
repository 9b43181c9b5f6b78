set -x
timeout 300 python tools/laguerre_time.py > gpurun_out/r2_lag_div.txt 2> gpurun_out/r2_lag_div_t.txt; cat gpurun_out/r2_lag_div_t.txt
timeout 300 python tools/hermite_time.py | tail -3
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_div.txt 2>&1
timeout 900 python -m pytest tests/test_laguerre_gpu.py tests/test_hermite_gpu.py tests/test_integrate_gpu.py -m gpu -q --timeout=600 2>&1 | tail -3
