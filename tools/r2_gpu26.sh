set -x
nvidia-smi -L
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x --timeout=200 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_torchrun_s2_2gpu.log 2>&1; echo torchrun rc=$?
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2_torchrun_s2_2gpu.log') if x.startswith('{')][-1]
d=json.loads(l)
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d.get('e2e',{}).get('value'), d.get('allgather'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -c 600
