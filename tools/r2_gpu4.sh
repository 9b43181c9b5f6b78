set -x
nvidia-smi -L
for v in r1 cur new new5 r1 new; do FGQ_LIB_PATH=$PWD/_exp/libfgq_$v.so python tools/time_tail.py 1e9 200 >> gpurun_out/r2_time_tail4.log 2>&1; done
FGQ_LIB_PATH=$PWD/_exp/libfgq_new.so python tools/time_tail.py 1000000001 100 >> gpurun_out/r2_time_tail4.log 2>&1
FGQ_LIB_PATH=$PWD/_exp/libfgq_r1.so python tools/time_tail.py 1000000001 100 >> gpurun_out/r2_time_tail4.log 2>&1
cat gpurun_out/r2_time_tail4.log
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest4.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest4.log; tail -12 gpurun_out/r2_pytest4.log | cut -c1-250
python tools/time_tail.py 1e9 3 > gpurun_out/r2_tail_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:legendre_asy_kernel -s 8 -c 2 -f -o gpurun_out/r2_legendre_1e9 python tools/time_tail.py 1e9 3 > gpurun_out/r2_ncu_full.log 2>&1; echo ncu rc=$?
ncu -i gpurun_out/r2_legendre_1e9.ncu-rep --page raw --csv > gpurun_out/r2_legendre_1e9_ncu_raw.csv 2>/dev/null
