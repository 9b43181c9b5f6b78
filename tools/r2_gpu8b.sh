set -x
NG=${1:-8}
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x --timeout=200 > gpurun_out/r2_pytest8b_multi_${NG}gpu.log 2>&1; echo rc=$? >> gpurun_out/r2_pytest8b_multi_${NG}gpu.log; tail -3 gpurun_out/r2_pytest8b_multi_${NG}gpu.log | cut -c1-250
for rep in 1 2; do
timeout 300 python bench.py --driver capi --gpus $NG --steps 20 --warmup 5 --no-e2e --no-allgather > gpurun_out/r2_capi8b_${NG}gpu_$rep.log 2>&1; echo capi rc=$?; tail -c 700 gpurun_out/r2_capi8b_${NG}gpu_$rep.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 --no-e2e --no-allgather > gpurun_out/r2_torchrun8b_${NG}gpu_$rep.log 2>&1; echo torchrun rc=$?; python -c "
import json;l=[x for x in open('gpurun_out/r2_torchrun8b_${NG}gpu_$rep.log') if x.startswith('{')][-1];d=json.loads(l);print('torchrun',d['value'],d['ms_per_step'])"
done
timeout 300 python bench.py --driver capi --gpus 4 --steps 20 --warmup 5 --no-e2e --no-allgather 2>&1 | tail -c 400
