set -x
timeout 300 python tools/laguerre_time.py > gpurun_out/r2_lag_tab.txt 2> gpurun_out/r2_lag_tab_t.txt; cat gpurun_out/r2_lag_tab_t.txt
timeout 600 python -m pytest tests/test_laguerre_gpu.py tests/test_small_batch_gpu.py -m gpu -q --timeout=600 2>&1 | tail -3
