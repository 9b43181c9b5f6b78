set -x
timeout 300 python tools/jacobi_time.py
timeout 900 python -m pytest tests/test_jacobi_gpu.py tests/test_fuzz_gpu.py -m gpu -q -s --timeout=600 2>&1 | grep -v "^\s*$" | tail -25 | cut -c1-300
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_lean.txt 2>&1; diff gpurun_out/r2_digest_lean.txt gpurun_out/r2_digest_2s.txt | head -20
