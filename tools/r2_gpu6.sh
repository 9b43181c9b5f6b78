set -x
nvidia-smi -L
python -m pytest tests/test_small_batch_gpu.py tests/test_laguerre_gpu.py tests/test_fuzz_gpu.py -m gpu -q -x > gpurun_out/r2_pytest6.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest6.log; tail -12 gpurun_out/r2_pytest6.log | cut -c1-250
python tools/small_batch_timing.py > gpurun_out/r2_small_batch_new.log 2>&1; cat gpurun_out/r2_small_batch_new.log
FGQ_LAGUERRE_BATCH=thread python tools/small_batch_timing.py > gpurun_out/r2_small_batch_old.log 2>&1; tail -3 gpurun_out/r2_small_batch_old.log
python tools/small_rule_digest.py > gpurun_out/r2_digest_new.txt 2>&1; FGQ_LAGUERRE_BATCH=thread python tools/small_rule_digest.py > gpurun_out/r2_digest_thread.txt 2>&1; diff gpurun_out/r2_digest_new.txt gpurun_out/r2_digest_thread.txt && echo DIGESTS_EQUAL
