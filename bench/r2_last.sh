set -u
export PYTHONPATH=. HVD_KERNEL_TIMEOUT_SECONDS=20
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_single.py -q 2>&1 | tail -3
timeout 200 python bench.py --parity --steps 5 --warmup 3 --no-extras > gpurun_out/bench1_parity.json 2> gpurun_out/bench1_parity.err; tail -c 900 gpurun_out/bench1_parity.json; tail -2 gpurun_out/bench1_parity.err
