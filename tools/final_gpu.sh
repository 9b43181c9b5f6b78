set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_final2.log 2>&1; echo rc=$? >> gpurun_out/pytest_final2.log; tail -3 gpurun_out/pytest_final2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final2.log 2>&1; tail -1 gpurun_out/smoke_final2.log
python bench.py > gpurun_out/bench_final2.log 2>&1; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final2.log 2>&1; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/cfg_launches_final2.csv python tools/run_configs.py 1 > gpurun_out/ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/small_batch_launches.csv python tools/run_small_batches.py > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:legendre_rec_tile -c 1 -f -o gpurun_out/final_rec_tile python tools/run_configs.py 1 > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobi_pass_kernel -c 6 -f -o gpurun_out/final_jacobi_pass python tools/run_configs.py 1 > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:small_batch -c 3 -f -o gpurun_out/final_small_batch python tools/run_small_batches.py > gpurun_out/ncu_e.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:laguerre_ -c 8 -f -o gpurun_out/final_laguerre python tools/run_configs.py 1 > gpurun_out/ncu_f.log 2>&1
for f in final_rec_tile final_jacobi_pass final_small_batch final_laguerre; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep
