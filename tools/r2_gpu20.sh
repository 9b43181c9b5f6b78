set -x
timeout 120 python tools/hermite_one.py || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:hermite_fused -s 2 -c 1 -o gpurun_out/r2_hermite_fused python tools/hermite_one.py > gpurun_out/r2_hermite_ncu.log 2>&1; tail -3 gpurun_out/r2_hermite_ncu.log
ls -la gpurun_out/
