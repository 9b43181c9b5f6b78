set -x
timeout 300 python tools/hermite_time.py
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_hdiv.txt 2>&1
timeout 900 python -m pytest tests/test_hermite_gpu.py -m gpu -q -s --timeout=600 2>&1 | grep -i "node err\|passed\|failed\|error" | tail -12
