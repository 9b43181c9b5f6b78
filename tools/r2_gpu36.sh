set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_last.txt 2>&1
timeout 600 python -m pytest tests/test_laguerre_gpu.py tests/test_hermite_gpu.py tests/test_small_batch_gpu.py -m gpu -q --timeout=600 2>&1 | tail -2
