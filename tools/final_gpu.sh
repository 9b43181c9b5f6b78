# End-of-round GPU run (one GPU): the whole -m gpu suite, smoke, the bench line, the reference arm on a short run, the
# full-range parity reports, the launch list of the bench command and the ncu capture of the headline kernel.
set -x
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r2_pytest_final.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_final.log; tail -4 gpurun_out/r2_pytest_final.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1; tail -1 gpurun_out/r2_smoke_final.log
timeout 900 python bench.py > gpurun_out/r2_bench_final.log 2>&1; echo bench rc=$?
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final_s20.log 2>&1; echo bench20 rc=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final.log 2>&1; echo ref rc=$?; tail -c 900 gpurun_out/r2_bench_ref_final.log
timeout 600 python tools/parity_report.py 1e9 gpurun_out/r2_parity_legendre_1e9.md > /dev/null 2>&1; tail -3 gpurun_out/r2_parity_legendre_1e9.md
timeout 600 python tools/parity_report.py 1000000001 gpurun_out/r2_parity_legendre_1e9p1.md > /dev/null 2>&1; tail -3 gpurun_out/r2_parity_legendre_1e9p1.md
timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/r2_bench_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs > gpurun_out/r2_bench_ncu.log 2>&1; echo launches rc=$?
timeout 300 python tools/jacobi_time.py > gpurun_out/r2_times_final.log 2>&1; timeout 300 python tools/hermite_time.py >> gpurun_out/r2_times_final.log 2>&1; timeout 300 python tools/laguerre_time.py 2>&1 | grep -v digest >> gpurun_out/r2_times_final.log; timeout 300 python tools/small_batch_timing.py >> gpurun_out/r2_times_final.log 2>&1; cat gpurun_out/r2_times_final.log | cut -c1-200
timeout 300 python tools/run_configs.py 1 > gpurun_out/r2_cfg_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_cfg_launches.csv python tools/run_configs.py 1 > gpurun_out/r2_cfg_ncu.log 2>&1; echo cfg rc=$?
timeout 300 python tools/r2_fp64_run.py > gpurun_out/r2_fp64_plain.log 2>&1 && timeout 900 ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_fp64_ncu.csv python tools/r2_fp64_run.py > gpurun_out/r2_fp64_ncu.log 2>&1; echo fp64 rc=$?
timeout 300 python tools/time_tail.py 1e9 3 > gpurun_out/r2_tail_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:legendre_asy_kernel -s 8 -c 2 -f -o gpurun_out/r2_legendre_1e9 python tools/time_tail.py 1e9 3 > gpurun_out/r2_ncu_full.log 2>&1; echo ncu rc=$?
ncu -i gpurun_out/r2_legendre_1e9.ncu-rep --page raw --csv > gpurun_out/r2_legendre_1e9_ncu_raw.csv 2>/dev/null
