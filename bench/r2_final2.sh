set -u
export PYTHONPATH=. HVD_KERNEL_TIMEOUT_SECONDS=30
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29599 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2_last.json 2> gpurun_out/bench2_last.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench2_last.json').read().strip().splitlines()[-1])
print('resnet50 2 GPUs', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['extra'].get('checks'), d['extra'].get('error'))
b = d['extra'].get('bert_large', {}); print('bert', b.get('value'), b.get('ms_per_step'), (b.get('e2e') or {}).get('ms_per_step'))
PY
wc -l gpurun_out/bench2_last.json; tail -2 gpurun_out/bench2_last.err
