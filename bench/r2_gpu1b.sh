# 1 GPU: kernel-only timings of the exchange kernel (simulation), GPU tests, ncu captures of every kernel family
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=. HVD_KERNEL_TIMEOUT_SECONDS=30
echo "== exchange kernel timing (simulation, device-timed)"
timeout 120 python - <<'PY'
import torch, time
from horovod_b200.ops import sim
for n in (1, 2, 4):
    for mb in (1, 16, 64):
        src = [torch.ones(mb << 20, device='cuda', dtype=torch.uint8) for _ in range(n)]
        dst = [torch.empty((mb << 20) * n, device='cuda', dtype=torch.uint8) for _ in range(n)]
        for ctas in (32, 128):
            for _ in range(2): sim.allgather(src, dst, ctas=ctas)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): sim.allgather(src, dst, ctas=ctas)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"allgather sim ranks={n} per-rank={mb} MiB ctas={ctas}: {ms*1e3:.1f} us  ({(mb << 20) * n * (1 + n) / n / (ms / 1e3) / 1e9:.0f} GB/s copied per rank)", flush=True)
PY
echo "== gpu tests (1 GPU)"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== ncu captures"
timeout 1500 bash bench/run_ncu_captures.sh 2>&1 | tail -30
