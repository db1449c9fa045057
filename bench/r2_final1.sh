set -u
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench1_final.json 2> gpurun_out/bench1_final.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench1_final.json').read().strip().splitlines()[-1])
print('resnet50 1 GPU', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
b = d['extra'].get('bert_large', {})
print('bert 1 GPU', b.get('value'), b.get('ms_per_step'), (b.get('e2e') or {}).get('value'), d['extra'].get('error'))
PY
tail -2 gpurun_out/bench1_final.err
timeout 120 python bench.py --impl reference
