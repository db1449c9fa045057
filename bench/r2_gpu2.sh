# round-2 validation on N GPUs (default 2): multi-GPU pytest, sanitizer re-run, allreduce sweeps (default vs NCCL), bench.py
set -u
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
export HVD_KERNEL_TIMEOUT_SECONDS=30 HOROVOD_LOG_LEVEL=warning PYTHONPATH=.
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multi-GPU pytest ($N GPUs)"
timeout 1200 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -60 > $OUT/r2_pytest_gpu_multi_${N}.log; tail -25 $OUT/r2_pytest_gpu_multi_${N}.log
echo "== racecheck (1 GPU simulation)"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 77 --launch-timeout 0 python tests/sanitizer_target.py 2>&1 | tail -12 > $OUT/r2_racecheck.log; tail -6 $OUT/r2_racecheck.log
echo "== allreduce sweep, plain tensors: default (pipelined on) vs pipelined off vs NCCL"
timeout 400 $TR --master-port 29531 bench/allreduce_sweep.py --sizes 4096,65536,1048576,16777216,67108864,268435456,1073741824 \
  --configs p2p:auto:128,p2p:auto:128+HVD_PIPELINED_ALLREDUCE=0,nccl --out $OUT/sweep${N}_plain.json 2>&1 | grep -v Warn | tail -40
echo "== allreduce sweep, registered tensors"
timeout 300 $TR --master-port 29532 bench/allreduce_sweep.py --symm --sizes 4096,65536,1048576,16777216,67108864,268435456,1073741824 \
  --configs p2p:auto:128 --out $OUT/sweep${N}_symm.json 2>&1 | grep -v Warn | tail -14
echo "== bench.py $N GPUs"
timeout 600 $TR --master-port 29533 bench.py --gpus $N --steps 15 --warmup 3 > $OUT/bench${N}.json 2> $OUT/bench${N}.err; tail -c 3000 $OUT/bench${N}.json; tail -5 $OUT/bench${N}.err
