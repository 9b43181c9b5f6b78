set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -x > gpurun_out/r2_pytest_s2.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest_s2.log; tail -6 gpurun_out/r2_pytest_s2.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_s2.log 2>&1; tail -2 gpurun_out/r2_smoke_s2.log
