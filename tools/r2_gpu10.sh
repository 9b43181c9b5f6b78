set -x
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_small_batch_gpu.py -m gpu -q --timeout=300 2>&1 | tail -4
timeout 300 python bench.py --driver capi --gpus 2 --steps 20 --warmup 5 --e2e-steps 3 2>&1 | tail -c 1500
