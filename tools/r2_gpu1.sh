set -x
nvidia-smi -L; nproc; free -g | head -2
python -m pytest tests/test_legendre_gpu.py tests/test_full_size_gpu.py -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest1.log; tail -5 gpurun_out/r2_pytest1.log
python bench.py > gpurun_out/r2_bench1.log 2>&1; echo bench rc=$?; tail -c 3000 gpurun_out/r2_bench1.log
python tools/host_call_timing.py > gpurun_out/r2_host_codec.log 2>&1; cat gpurun_out/r2_host_codec.log
FGQ_CODEC_FUSED=1 python tools/host_call_timing.py > gpurun_out/r2_host_codec_fused.log 2>&1; cat gpurun_out/r2_host_codec_fused.log
python -m pytest tests -m gpu -q --deselect tests/test_legendre_gpu.py --deselect tests/test_full_size_gpu.py > gpurun_out/r2_pytest1b.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest1b.log; tail -5 gpurun_out/r2_pytest1b.log
