set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_small_batch_gpu.py tests/test_laguerre_gpu.py tests/test_fuzz_gpu.py -m gpu -q -x --timeout=300 > gpurun_out/r2_pytest7b.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest7b.log; tail -6 gpurun_out/r2_pytest7b.log | cut -c1-250
timeout 300 python tools/small_batch_timing.py 2>&1 | tail -1
FGQ_LAGUERRE_BATCH=thread timeout 300 python tools/small_batch_timing.py 2>&1 | tail -1
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_new.txt 2>&1; FGQ_LAGUERRE_BATCH=thread timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_thread.txt 2>&1; diff gpurun_out/r2_digest_new.txt gpurun_out/r2_digest_thread.txt && echo DIGESTS_EQUAL
timeout 600 python tools/jacobi_c3_diag.py 1e7
FGQ_LIB_PATH=$PWD/_exp/libfgq_norot.so timeout 600 python tools/jacobi_c3_diag.py 1e7
timeout 600 python tools/jacobi_c3_diag.py 1e6
FGQ_LIB_PATH=$PWD/_exp/libfgq_norot.so timeout 600 python tools/jacobi_c3_diag.py 1e6
