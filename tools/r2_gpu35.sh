set -x
timeout 120 python tools/laguerre_one.py || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:laguerre_ -s 12 -c 6 -o gpurun_out/r2_laguerre python tools/laguerre_one.py > gpurun_out/r2_laguerre_ncu.log 2>&1; tail -3 gpurun_out/r2_laguerre_ncu.log
