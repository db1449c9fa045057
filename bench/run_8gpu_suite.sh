#!/usr/bin/env bash
# The 8-GPU measurement suite (one gpurun call): headline training benches for every BASELINE.json config, the NCCL-backend
# arm of the same engine, the allreduce sweeps and the hierarchical-allreduce check.  Results land in gpurun_out/.
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
export HVD_KERNEL_TIMEOUT_SECONDS=30 HOROVOD_LOG_LEVEL=warning
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # name timeout cmd...
  local name=$1 t=$2; shift 2
  local t0=$(date +%s)
  timeout $t "$@" > $OUT/$name.json 2> $OUT/$name.err
  echo "[$name] rc=$? $(( $(date +%s) - t0 ))s $(tail -c 300 $OUT/$name.err | tr '\n' ' ' | cut -c1-200)"
  head -c 700 $OUT/$name.json; echo
}
run resnet50_${N}gpu 150 $TR --master-port 29501 bench.py --gpus $N --steps 30 --warmup 5
run resnet50_${N}gpu_nccl 150 env HVD_GPU_BACKEND=nccl $TR --master-port 29502 bench.py --gpus $N --steps 30 --warmup 5
run bert_large_${N}gpu 200 $TR --master-port 29503 bench.py --gpus $N --model bert-large --steps 12 --warmup 3
run gpt2_medium_adasum_${N}gpu 200 $TR --master-port 29504 bench.py --gpus $N --model gpt2-medium --op adasum --steps 8 --warmup 3
timeout 150 $TR --master-port 29505 bench/allreduce_sweep.py --symm --sizes 16777216,134217728,1073741824 \
  --configs p2p:auto:128,p2p:auto:256:8,p2p:auto:256:8:131072 --out $OUT/sweep${N}_symm_tuned.json 2>&1 | grep -v Warn | tail -24
if [ "${HVD_SUITE_FULL:-0}" = "1" ]; then
timeout 150 $TR --master-port 29506 bench/allreduce_sweep.py --sizes 16777216,134217728,1073741824 \
  --configs p2p:auto:128,nccl --out $OUT/sweep${N}_plain_tuned.json 2>&1 | grep -v Warn | tail -12
timeout 200 python -m pytest tests/test_gpu_multi.py -q -x -k "hierarchical" 2>&1 | tail -4
fi
