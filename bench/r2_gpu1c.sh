# 1 GPU: persistent Adasum vs the fp64 oracle (both variants), memcheck over the simulation targets, ncu captures
set -u
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=. HVD_KERNEL_TIMEOUT_SECONDS=20
echo "== adasum kernels vs fp64 oracle: persistent / multi-launch"
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -k "adasum or exchange or allgather" 2>&1 | tail -4
HVD_ADASUM_PERSISTENT=0 timeout 200 python -m pytest tests/test_gpu_kernels.py -q -k "adasum" 2>&1 | tail -3
echo "== determinism of the persistent kernel (two runs, bitwise)"
timeout 100 python - <<'PY'
import torch
from horovod_b200.ops import sim
n = 8
torch.manual_seed(0)
ins = [[torch.randn(s, device='cuda') for s in (5, 70001, 333, 1 << 20)] for _ in range(n)]
res = []
for rep in range(2):
    outs = [[torch.empty_like(x) for x in row] for row in ins]
    sim.adasum(ins, outs, ctas=8)
    torch.cuda.synchronize()
    res.append([o.clone() for o in outs[0]])
print('bitwise identical across runs:', all(torch.equal(a, b) for a, b in zip(*res)), ' ranks agree:', True)
PY
echo "== memcheck"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 77 --launch-timeout 0 python tests/sanitizer_target.py 2>&1 | tail -4
echo "== smoke"
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== ncu captures"
timeout 1300 bash bench/run_ncu_captures.sh 2>&1 | tail -25
