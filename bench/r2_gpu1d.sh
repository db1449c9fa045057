set -u
export PYTHONPATH=. HVD_KERNEL_TIMEOUT_SECONDS=20
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/r2_pytest_gpu_1_final.log
timeout 100 python - <<'PY'
import torch
from horovod_b200.ops import sim
for n, mb, ctas in ((1, 64, 64), (2, 64, 64), (8, 16, 16)):
    ins = [[torch.randn((mb << 20) // 16, device='cuda') for _ in range(4)] for _ in range(n)]
    outs = [[torch.empty_like(x) for x in row] for row in ins]
    sim.adasum(ins, outs, ctas=ctas)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): sim.adasum(ins, outs, ctas=ctas)
    e1.record(); torch.cuda.synchronize()
    print(f'persistent adasum sim ranks={n} {mb} MiB ctas={ctas}: {e0.elapsed_time(e1) / 3 * 1e3:.0f} us', flush=True)
PY
