set -x
nvidia-smi -L; nproc
python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_pytest2_multi.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest2_multi.log; tail -15 gpurun_out/r2_pytest2_multi.log
python -m pytest tests/test_legendre_gpu.py tests/test_full_size_gpu.py "tests/test_jacobi_gpu.py::test_config_c3_every_node_against_the_oracle" -m gpu -x -q -s > gpurun_out/r2_pytest2.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest2.log; tail -8 gpurun_out/r2_pytest2.log; grep "C3 full" gpurun_out/r2_pytest2.log
python tools/time_tail.py 1e9 100 > gpurun_out/r2_time_tail.log 2>&1
FGQ_LIB_PATH=$PWD/_exp/libfgq_mb5.so python tools/time_tail.py 1e9 100 >> gpurun_out/r2_time_tail.log 2>&1
FGQ_LIB_PATH=$PWD/_exp/libfgq_mb6.so python tools/time_tail.py 1e9 100 >> gpurun_out/r2_time_tail.log 2>&1
python tools/time_tail.py 1e9 100 >> gpurun_out/r2_time_tail.log 2>&1
cat gpurun_out/r2_time_tail.log
python bench.py > gpurun_out/r2_bench2.log 2>&1; echo bench rc=$?; tail -c 1500 gpurun_out/r2_bench2.log
python bench.py --driver capi --gpus 1 --steps 20 > gpurun_out/r2_bench2_capi.log 2>&1; echo capi rc=$?; tail -c 1500 gpurun_out/r2_bench2_capi.log
python tools/r2_fp64_run.py > gpurun_out/r2_fp64_plain.log 2>&1 && ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_fp64_ncu.csv python tools/r2_fp64_run.py > gpurun_out/r2_fp64_ncu.log 2>&1; echo fp64 rc=$?
python tools/time_tail.py 1e9 3 > gpurun_out/r2_tail_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:legendre_asy_kernel -s 8 -c 2 -f -o gpurun_out/r2_legendre_1e9 python tools/time_tail.py 1e9 3 > gpurun_out/r2_ncu_full.log 2>&1; echo ncu rc=$?
ncu -i gpurun_out/r2_legendre_1e9.ncu-rep --page raw --csv > gpurun_out/r2_legendre_1e9_ncu_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
