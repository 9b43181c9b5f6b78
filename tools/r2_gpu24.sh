set -x
timeout 300 python tools/run_configs.py 1 > gpurun_out/r2_cfg_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_cfg_launches.csv python tools/run_configs.py 1 > gpurun_out/r2_cfg_ncu.log 2>&1; echo cfg rc=$?
timeout 300 python tools/r2_fp64_run.py > gpurun_out/r2_fp64_plain.log 2>&1 && timeout 900 ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_fp64_ncu.csv python tools/r2_fp64_run.py > gpurun_out/r2_fp64_ncu.log 2>&1; echo fp64 rc=$?
ls -la gpurun_out
