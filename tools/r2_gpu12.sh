set -x
timeout 300 python tools/jacobi_time.py
FGQ_JACOBI_TWO_STREAMS=0 timeout 300 python tools/jacobi_time.py
timeout 900 python -m pytest tests/test_jacobi_gpu.py tests/test_fuzz_gpu.py tests/test_concurrency_gpu.py tests/test_small_batch_gpu.py -m gpu -q --timeout=300 2>&1 | tail -4
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_2s.txt 2>&1; diff gpurun_out/r2_digest_2s.txt gpurun_out/r2_digest_nopdl.txt && echo DIGESTS_EQUAL
