#!/usr/bin/env bash
# Everything written after round 1's GPU budget was spent, in the order it should be validated (cheapest first).
#   1 GPU :  bash bench/run_pending_gpu_validation.sh 1
#   N GPUs:  bash bench/run_pending_gpu_validation.sh 8      (also runs the 1-GPU part)
set -u
N=${1:-1}
OUT=gpurun_out
mkdir -p $OUT
export HVD_KERNEL_TIMEOUT_SECONDS=30 HOROVOD_LOG_LEVEL=warning HVD_RUN_NEW_GPU_TESTS=1
echo "== [1 GPU] gated kernel tests (TMA exchange variant in the single-GPU simulation)"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "tma" 2>&1 | tail -5
echo "== [1 GPU] software-pipelined allreduce (P2P reduce stage) in the simulation"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "pipelined" 2>&1 | tail -5
echo "== [1 GPU] compute-sanitizer memcheck over the P2P kernels"
HVD_RUN_COMPUTE_SANITIZER=1 timeout 900 python -m pytest tests/test_sanitizers.py -q -x -m gpu 2>&1 | tail -5
if [ "$N" -ge 2 ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  echo "== [$N GPUs] gated multi-GPU tests (extra op matrix on CUDA, hierarchical allreduce with 2 fake hosts)"
  timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -8
  echo "== [$N GPUs] small-message latency: shared-memory control plane (now really in the negotiation path) vs TCP"
  for cp in auto tcp; do
    HVD_CONTROL_PLANE=$cp timeout 150 $TR --master-port 29520 bench/allreduce_sweep.py --sizes 4096,65536,1048576 --configs p2p:auto:128 \
      --out $OUT/sweep${N}_latency_cp_${cp}.json 2>&1 | grep -v Warn | tail -6
  done
  echo "== [$N GPUs] allgather / broadcast / alltoall / reducescatter vs torch.distributed NCCL"
  timeout 200 $TR --master-port 29521 bench/collective_sweep.py --nccl --out $OUT/collectives${N}.json 2>&1 | grep -v Warn | tail -30
  echo "== [$N GPUs] same, TMA bulk-copy exchange kernel"
  HVD_EXCHANGE_TMA=1 timeout 200 $TR --master-port 29522 bench/collective_sweep.py --ops allgather,broadcast,alltoall --tag tma \
    --out $OUT/collectives${N}_tma.json 2>&1 | grep -v Warn | tail -20
  echo "== [$N GPUs] software-pipelined allreduce of plain tensors vs the three-phase kernel and NCCL"
  HVD_PIPELINED_ALLREDUCE=1 timeout 150 $TR --master-port 29524 bench/allreduce_sweep.py --sizes 33554432,134217728,1073741824 \
    --configs p2p:auto:128,nccl --out $OUT/sweep${N}_plain_pipelined.json 2>&1 | grep -v Warn | tail -12
  echo "== [$N GPUs] pipelined small-message allreduce (8 in flight)"
  timeout 150 $TR --master-port 29523 bench/allreduce_sweep.py --inflight 8 --sizes 4096,65536,1048576 --configs p2p:auto:128,nccl \
    --out $OUT/sweep${N}_inflight8.json 2>&1 | grep -v Warn | tail -12
fi
