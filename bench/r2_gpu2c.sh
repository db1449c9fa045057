# 2 GPUs: multi-GPU pytest (elastic on GPUs, latency lane, IPC, wire dtype), collectives vs NCCL, plain sweep, bench
set -u
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
export HVD_KERNEL_TIMEOUT_SECONDS=30 HOROVOD_LOG_LEVEL=warning PYTHONPATH=. HVD_CACHE_DIR=/tmp/hvdcache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multi-GPU pytest ($N GPUs)"
timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -45 > $OUT/r2_pytest_gpu_multi_${N}c.log; tail -8 $OUT/r2_pytest_gpu_multi_${N}c.log
echo "== other collectives vs NCCL"
timeout 150 $TR --master-port 29572 bench/collective_sweep.py --nccl --out $OUT/collectives${N}.json 2>&1 | grep "^==\| B " | tail -30
echo "== allreduce sweep plain: default | nccl"
HOROVOD_LOG_LEVEL=info timeout 300 $TR --master-port 29571 bench/allreduce_sweep.py --sizes 4096,65536,262144,1048576,4194304,16777216,67108864,1073741824 \
  --configs p2p:auto:128,nccl --out $OUT/sweep${N}_plain_c.json 2>&1 | grep "calibration\|^==\| B " | tail -30
echo "== bench.py $N GPUs"
timeout 400 $TR --master-port 29573 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench${N}c.json 2> $OUT/bench${N}c.err; python - <<PY
import json
d = json.loads(open("$OUT/bench${N}c.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["extra"].get("checks"), d["extra"].get("error"))
print("bert", d["extra"].get("bert_large", {}).get("value"), d["extra"].get("bert_large", {}).get("ms_per_step"))
for r in d["extra"].get("allreduce_busbw", {}).get("rows", []): print(r)
PY
wc -l $OUT/bench${N}c.json; tail -3 $OUT/bench${N}c.err
