set -x
timeout 300 python tools/laguerre_time.py > gpurun_out/r2_lag_fast.txt 2> gpurun_out/r2_lag_fast_t.txt; cat gpurun_out/r2_lag_fast_t.txt
FGQ_LAGUERRE_FAST=0 timeout 300 python tools/laguerre_time.py > gpurun_out/r2_lag_seq.txt 2> gpurun_out/r2_lag_seq_t.txt; cat gpurun_out/r2_lag_seq_t.txt
FGQ_LAGUERRE_WINDOW=64 timeout 300 python tools/laguerre_time.py > gpurun_out/r2_lag_fb.txt 2> gpurun_out/r2_lag_fb_t.txt; cat gpurun_out/r2_lag_fb_t.txt
cat gpurun_out/r2_lag_fast.txt
diff gpurun_out/r2_lag_fast.txt gpurun_out/r2_lag_seq.txt && echo FAST_EQ_SEQ
diff gpurun_out/r2_lag_fb.txt gpurun_out/r2_lag_seq.txt && echo FALLBACK_EQ_SEQ
timeout 600 python -m pytest tests/test_laguerre_gpu.py -m gpu -q --timeout=300 2>&1 | tail -4
