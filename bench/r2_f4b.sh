set -u
export HVD_KERNEL_TIMEOUT_SECONDS=30 HOROVOD_LOG_LEVEL=warning PYTHONPATH=. HVD_CACHE_DIR=/tmp/hvdcache
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29591 bench/allreduce_sweep.py \
  --sizes 4194304,16777216,67108864,268435456,1073741824 --configs p2p:auto:128+HVD_IPC_MAX_RANKS=4 --out gpurun_out/sweep4_plain_ipc.json 2>&1 | grep "^==\| B " | tail -12
