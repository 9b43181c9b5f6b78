set -x
timeout 600 python -m pytest tests/test_small_batch_gpu.py tests/test_laguerre_gpu.py tests/test_concurrency_gpu.py -m gpu -q --timeout=300 > gpurun_out/r2_pytest7d.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest7d.log; tail -4 gpurun_out/r2_pytest7d.log | cut -c1-250
timeout 300 python tools/small_batch_timing.py 2>&1 | tail -1
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_new.txt 2>&1; FGQ_LAGUERRE_BATCH=thread timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_thread.txt 2>&1; diff gpurun_out/r2_digest_new.txt gpurun_out/r2_digest_thread.txt && echo DIGESTS_EQUAL
