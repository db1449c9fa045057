# final measurements, part 1 (N GPUs): plain-tensor allreduce sweep (default | dual lane off | NCCL data plane) + the other
# collectives vs torch.distributed NCCL
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
export HVD_KERNEL_TIMEOUT_SECONDS=30 HOROVOD_LOG_LEVEL=warning PYTHONPATH=. HVD_CACHE_DIR=/tmp/hvdcache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== allreduce sweep $N GPUs, plain tensors"
HOROVOD_LOG_LEVEL=info timeout 240 $TR --master-port 29581 bench/allreduce_sweep.py --sizes 4096,65536,1048576,16777216,67108864,268435456,1073741824 \
  --configs ${2:-p2p:auto:128,p2p:auto:128+HVD_DUAL_LANE_ALLREDUCE=0,nccl} --out $OUT/sweep${N}_plain_final.json 2>&1 | grep "calibration\|^==\| B " | tail -40
echo "== other collectives vs NCCL"
timeout 120 $TR --master-port 29582 bench/collective_sweep.py --nccl --out $OUT/collectives${N}.json 2>&1 | grep "^==\| B " | tail -30
