set -x
timeout 300 python tools/jacobi_time.py
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_jnorm.txt 2>&1
timeout 900 python -m pytest tests/test_jacobi_gpu.py tests/test_fuzz_gpu.py -m gpu -q -s --timeout=600 2>&1 | grep -i "C3 full\|passed\|failed\|error" | tail -8
