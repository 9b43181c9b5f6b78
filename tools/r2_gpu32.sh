set -x
timeout 600 ncu --set full --import-source on --clock-control none -k regex:jacobi_pass_kernel -c 6 -o gpurun_out/r2_jacobi_pass python tools/jacobi_one.py > gpurun_out/r2_jacobi_ncu.log 2>&1; tail -3 gpurun_out/r2_jacobi_ncu.log
