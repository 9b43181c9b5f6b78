set -x
nvidia-smi -L; nproc
NG=${1:-2}
python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/r2_pytest5_multi_${NG}gpu.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest5_multi_${NG}gpu.log; tail -12 gpurun_out/r2_pytest5_multi_${NG}gpu.log | cut -c1-250
for g in 1 2 4 8; do
  if [ $g -le $NG ]; then
    python bench.py --driver capi --gpus $g --steps 20 --warmup 5 > gpurun_out/r2_capi_${g}gpu.log 2>&1; echo capi $g rc=$?; tail -c 2500 gpurun_out/r2_capi_${g}gpu.log
  fi
done
for g in 2 4 8; do
  if [ $g -le $NG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 20 --warmup 5 > gpurun_out/r2_torchrun_${g}gpu.log 2>&1; echo torchrun $g rc=$?; tail -c 2500 gpurun_out/r2_torchrun_${g}gpu.log
  fi
done
FGQ_LIB_PATH=$PWD/_exp/libfgq_new2.so python tools/time_tail.py 1000000001 100
FGQ_LIB_PATH=$PWD/_exp/libfgq_new2.so python tools/time_tail.py 1e9 100
python -m pytest tests/test_legendre_gpu.py tests/test_full_size_gpu.py -m gpu -q -x > gpurun_out/r2_pytest5.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest5.log; tail -5 gpurun_out/r2_pytest5.log | cut -c1-250
