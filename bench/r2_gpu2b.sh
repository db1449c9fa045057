# round-2 second validation on 2 GPUs: IPC registration, dual-lane large allreduce, calibration, collectives vs NCCL
set -u
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
export HVD_KERNEL_TIMEOUT_SECONDS=30 HOROVOD_LOG_LEVEL=warning PYTHONPATH=. HVD_CACHE_DIR=/tmp/hvdcache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='grep -v Warn\|^\*\|OMP_NUM\|^$'
echo "== stress: sanitizer target x20 without a tool"
for i in $(seq 1 20); do timeout 60 python tests/sanitizer_target.py > $OUT/stress.log 2>&1 || { echo "iteration $i FAILED"; tail -3 $OUT/stress.log; break; }; done; echo "stress done"
echo "== multi-GPU pytest ($N GPUs)"
timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -40 > $OUT/r2_pytest_gpu_multi_${N}b.log; tail -8 $OUT/r2_pytest_gpu_multi_${N}b.log
echo "== allreduce sweep plain: default (IPC + dual lane + calibration) | no IPC | no IPC, no dual lane | nccl"
HOROVOD_LOG_LEVEL=info timeout 500 $TR --master-port 29561 bench/allreduce_sweep.py --sizes 4096,65536,1048576,4194304,16777216,67108864,268435456,1073741824 \
  --configs p2p:auto:128,p2p:auto:128+HVD_IPC_REGISTRATION=0,p2p:auto:128+HVD_IPC_REGISTRATION=0+HVD_DUAL_LANE_ALLREDUCE=0,nccl --out $OUT/sweep${N}_plain_b.json 2>&1 | grep "calibration\|^==\|  B  \| B " | tail -50
echo "== other collectives vs NCCL"
timeout 150 $TR --master-port 29562 bench/collective_sweep.py --nccl --out $OUT/collectives${N}.json 2>&1 | grep "^==\| B " | tail -30
echo "== bench.py $N GPUs"
timeout 400 $TR --master-port 29563 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench${N}b.json 2> $OUT/bench${N}b.err; python - <<PY
import json
d = json.load(open("$OUT/bench${N}b.json"))
print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["extra"].get("checks"), d["extra"].get("error"))
print("bert", d["extra"].get("bert_large", {}).get("value"), d["extra"].get("bert_large", {}).get("ms_per_step"))
for r in d["extra"].get("allreduce_busbw", {}).get("rows", []): print(r)
PY
tail -3 $OUT/bench${N}b.err
