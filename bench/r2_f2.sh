# final measurements, part 2 (N GPUs): bench.py (ResNet-50 + extras: checks, busBW vs NCCL, BERT-large) and GPT-2 medium
# with the GPU Adasum kernel
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
export HVD_KERNEL_TIMEOUT_SECONDS=30 HOROVOD_LOG_LEVEL=warning PYTHONPATH=. HVD_CACHE_DIR=/tmp/hvdcache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== bench.py $N GPUs"
timeout 400 $TR --master-port 29583 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench${N}_final.json 2> $OUT/bench${N}_final.err
python - <<PY
import json
d = json.loads(open("$OUT/bench${N}_final.json").read().strip().splitlines()[-1])
print("resnet50", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "in_graph", d["config"]["allreduce_in_graph"])
print("checks", d["extra"].get("checks"), d["extra"].get("error"))
b = d["extra"].get("bert_large", {})
print("bert", b.get("value"), b.get("ms_per_step"), "e2e", (b.get("e2e") or {}).get("value"), b.get("config"))
for r in d["extra"].get("allreduce_busbw", {}).get("rows", []): print(r)
PY
tail -2 $OUT/bench${N}_final.err
echo "== GPT-2 medium, op=Adasum"
timeout 300 $TR --master-port 29584 bench.py --gpus $N --model gpt2-medium --op adasum --steps 5 --warmup 3 --no-extras > $OUT/gpt2_adasum${N}.json 2> $OUT/gpt2_adasum${N}.err
python - <<PY
import json
d = json.loads(open("$OUT/gpt2_adasum${N}.json").read().strip().splitlines()[-1])
print("gpt2-medium adasum", d["value"], d["unit"], d["ms_per_step"], "ms/step  launches/step", d["gpu_launches"] / d["steps"], d["config"].get("cuda_graph"), d["config"].get("cuda_graph_fallback"))
PY
tail -2 $OUT/gpt2_adasum${N}.err
