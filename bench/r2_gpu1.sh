set -u
mkdir -p gpurun_out
export HVD_KERNEL_TIMEOUT_SECONDS=30 HOROVOD_LOG_LEVEL=warning
nvidia-smi -L > gpurun_out/r2_gpus.txt
echo "== gated+ungated gpu tests"
HVD_RUN_NEW_GPU_TESTS=1 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_pytest_gpu_1.log; tail -15 gpurun_out/r2_pytest_gpu_1.log
echo "== memcheck"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 77 --launch-timeout 0 python tests/sanitizer_target.py 2>&1 | tail -25 > gpurun_out/r2_memcheck.log; tail -6 gpurun_out/r2_memcheck.log
echo "== racecheck"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 77 --launch-timeout 0 python tests/sanitizer_target.py 2>&1 | tail -40 > gpurun_out/r2_racecheck.log; tail -8 gpurun_out/r2_racecheck.log
echo "== synccheck"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 77 --launch-timeout 0 python tests/sanitizer_target.py 2>&1 | tail -25 > gpurun_out/r2_synccheck.log; tail -6 gpurun_out/r2_synccheck.log
