set -x
timeout 300 python tools/small_batch_timing.py 2>&1 | tail -3 | cut -c1-250
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_pipe.txt 2>&1
timeout 900 python -m pytest tests/test_small_batch_gpu.py tests/test_fuzz_gpu.py tests/test_concurrency_gpu.py -m gpu -q --timeout=600 2>&1 | tail -3
