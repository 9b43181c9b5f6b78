set -x
nvidia-smi -L
python -m pytest tests/test_multi_gpu.py tests/test_integrate_gpu.py tests/test_laguerre_gpu.py -m gpu -q > gpurun_out/r2_pytest3.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest3.log; tail -40 gpurun_out/r2_pytest3.log | cut -c1-250
for v in r1 cur mb5 mb6 noint r1 cur; do FGQ_LIB_PATH=$PWD/_exp/libfgq_$v.so python tools/time_tail.py 1e9 200 >> gpurun_out/r2_time_tail3.log 2>&1; done
cat gpurun_out/r2_time_tail3.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.sw_thermal_slowdown --format=csv -lms 100 > gpurun_out/r2_clocks3.csv &
SMI=$!
FGQ_LIB_PATH=$PWD/_exp/libfgq_cur.so python tools/time_tail.py 1e9 400 >> gpurun_out/r2_time_tail3.log 2>&1
FGQ_LIB_PATH=$PWD/_exp/libfgq_r1.so python tools/time_tail.py 1e9 400 >> gpurun_out/r2_time_tail3.log 2>&1
kill $SMI
tail -2 gpurun_out/r2_time_tail3.log
awk -F, 'NR>1{print $2, $4, $6, $7}' gpurun_out/r2_clocks3.csv | sort | uniq -c | sort -rn | head -20
