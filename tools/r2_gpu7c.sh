set -x
timeout 900 python -m pytest tests/test_small_batch_gpu.py tests/test_laguerre_gpu.py tests/test_fuzz_gpu.py tests/test_jacobi_gpu.py -m gpu -q --timeout=300 > gpurun_out/r2_pytest7c.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest7c.log; tail -8 gpurun_out/r2_pytest7c.log | cut -c1-250; grep "C3 full" gpurun_out/r2_pytest7c.log
timeout 300 python tools/small_batch_timing.py 2>&1 | tail -3
timeout 600 python tools/jacobi_c3_diag.py 1e7
timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_new.txt 2>&1; FGQ_LAGUERRE_BATCH=thread timeout 300 python tools/small_rule_digest.py > gpurun_out/r2_digest_thread.txt 2>&1; diff gpurun_out/r2_digest_new.txt gpurun_out/r2_digest_thread.txt && echo DIGESTS_EQUAL
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench7c.log 2>&1; echo bench rc=$?
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_bench7c.log') if x.startswith('{')][-1]
d=json.loads(l)
print(d['value'], d['ms_per_step'])
for k,v in d['other_configs'].items(): print(k, v.get('device_ms') or v.get('stream_ms_incl_host_planning'))
PY
