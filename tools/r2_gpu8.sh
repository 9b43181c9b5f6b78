# lean multi-GPU run (charged N x box time): the multi-GPU C ABI tests, then the headline metric at 1/2/4/N through
# the single-process C ABI and through torchrun, with the all-gather timings
set -x
NG=${1:-8}
nvidia-smi -L | wc -l; nproc; free -g | head -2
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x --timeout=300 > gpurun_out/r2_pytest8_multi_${NG}gpu.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest8_multi_${NG}gpu.log; tail -5 gpurun_out/r2_pytest8_multi_${NG}gpu.log | cut -c1-250
for g in 1 2 4 8; do
  if [ $g -le $NG ]; then
    timeout 600 python bench.py --driver capi --gpus $g --steps 20 --warmup 5 --e2e-steps 5 > gpurun_out/r2_capi8_${g}gpu.log 2>&1; echo capi $g rc=$?; tail -c 2600 gpurun_out/r2_capi8_${g}gpu.log
  fi
done
for g in 2 4 8; do
  if [ $g -le $NG ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 20 --warmup 5 --e2e-steps 5 > gpurun_out/r2_torchrun8_${g}gpu.log 2>&1; echo torchrun $g rc=$?; tail -c 2600 gpurun_out/r2_torchrun8_${g}gpu.log
  fi
done
