# round-2 measurements on N GPUs (default 8): sweeps vs NCCL, bench.py (ResNet-50 + extras), BERT-large tuning, multi-GPU tests
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
export HVD_KERNEL_TIMEOUT_SECONDS=30 HOROVOD_LOG_LEVEL=warning PYTHONPATH=.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
SIZES=4096,65536,1048576,16777216,67108864,268435456,1073741824
echo "== allreduce sweep $N GPUs, plain tensors"
timeout 420 $TR --master-port 29541 bench/allreduce_sweep.py --sizes $SIZES \
  --configs p2p:auto:128,p2p:auto:128+HVD_PIPELINED_ALLREDUCE=0,nccl --out $OUT/sweep${N}_plain.json 2>&1 | grep -v "Warn\|^\*\|OMP_NUM\|^$" | tail -40
echo "== pipelined chunk size"
timeout 300 $TR --master-port 29542 bench/allreduce_sweep.py --sizes 67108864,268435456,1073741824 \
  --configs p2p:auto:128+HVD_PIPE_CHUNK_BYTES=2097152,p2p:auto:128+HVD_PIPE_CHUNK_BYTES=8388608,p2p:auto:128+HVD_PIPE_CHUNK_BYTES=16777216 --out $OUT/sweep${N}_pipe_chunk.json 2>&1 | grep -v "Warn\|^\*\|OMP_NUM\|^$" | tail -20
echo "== allreduce sweep $N GPUs, registered tensors"
timeout 300 $TR --master-port 29543 bench/allreduce_sweep.py --symm --sizes $SIZES --configs p2p:auto:128 --out $OUT/sweep${N}_symm.json 2>&1 | grep -v "Warn\|^\*\|OMP_NUM\|^$" | tail -12
echo "== other collectives vs NCCL"
timeout 300 $TR --master-port 29544 bench/collective_sweep.py --nccl --out $OUT/collectives${N}.json 2>&1 | grep -v "Warn\|^\*\|OMP_NUM\|^$" | tail -40
echo "== bench.py $N GPUs (ResNet-50 + extras)"
timeout 600 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/bench${N}.json 2> $OUT/bench${N}.err; tail -c 4500 $OUT/bench${N}.json; tail -3 $OUT/bench${N}.err
echo "== BERT-large tuning"
for cfg in "16 32" "64 32" "32 128" "32 32 bf16"; do
  set -- $cfg
  extra=""; [ "${3:-}" = "bf16" ] && extra="--bucket-wire-dtype bf16"
  HVD_GRAPH_COMM_CTAS=$1 timeout 300 $TR --master-port 29546 bench.py --gpus $N --model bert-large --steps 10 --warmup 3 --no-extras --bucket-cap-mb $2 $extra \
    > $OUT/bert${N}_ctas$1_cap$2${3:-}.json 2>> $OUT/bench${N}.err
  python - <<PY
import json
d = json.load(open("$OUT/bert${N}_ctas$1_cap$2${3:-}.json"))
print("bert-large ctas=$1 cap=$2 ${3:-fp32}:", d["value"], d["unit"], d["ms_per_step"], "ms/step e2e", d["e2e"]["value"], "in_graph", d["config"]["allreduce_in_graph"])
PY
done
echo "== multi-GPU pytest ($N GPUs)"
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "graph_nodes or fused_collectives or hierarchical or ops_matrix_p2p or new_checks" 2>&1 | tail -40 > $OUT/r2_pytest_gpu_multi_${N}.log; tail -12 $OUT/r2_pytest_gpu_multi_${N}.log
