"""broadcast_optimizer_state for EVERY optimizer class torch ships (reference: test_torch.py::test_broadcast_state /
test_broadcast_state_options): ranks start from different hyper-parameters and different state, afterwards all equal root's.
Also: state present on the root only (restored from a checkpoint), and wrapped (hvd.DistributedOptimizer) optimizers."""
import inspect

import torch

import horovod_b200.torch as hvd

hvd.init()
rank, size = hvd.rank(), hvd.size()
ROOT = size - 1                      # not rank 0 on purpose


def classes():
    out = []
    for name, cls in sorted(vars(torch.optim).items()):
        if inspect.isclass(cls) and issubclass(cls, torch.optim.Optimizer) and cls is not torch.optim.Optimizer:
            if name in ('LBFGS', 'SparseAdam'):          # closure-based / sparse-gradient only
                continue
            out.append(cls)
    return out


def make_model(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(5, 7, bias=False), torch.nn.Tanh(), torch.nn.Linear(7, 3, bias=False))  # 2-D only: Muon


def perturb(defaults, r):
    """Rank-dependent values for every numeric hyper-parameter the constructor exposes."""
    kw = {}
    for k, v in defaults.items():
        if k in ('differentiable', 'foreach', 'fused', 'capturable', 'maximize', 'nesterov', 'amsgrad', 'centered', 'decoupled_weight_decay'):
            continue
        if isinstance(v, bool) or v is None:
            continue
        if isinstance(v, float):
            if 0.5 <= v < 1:                        # decay rates (alpha, rho, ...) must stay below 1
                kw[k] = v * (1 - 0.01 * r)
            else:
                kw[k] = v * (1 + 0.1 * r) if v else (0.01 * r if k in ('weight_decay', 'momentum') else v)
        elif isinstance(v, (tuple, list)) and all(isinstance(x, float) for x in v):
            kw[k] = tuple(min(x * (1 - 0.01 * r), 0.9999) if x < 1 else x * (1 + 0.01 * r) for x in v)
    return kw


def flat_state(opt):
    sd = opt.state_dict()
    vals = []
    for pid in sorted(sd['state'], key=str):
        for k in sorted(sd['state'][pid]):
            v = sd['state'][pid][k]
            vals.append((pid, k, v.detach().double().flatten().tolist() if torch.is_tensor(v) else v))
    groups = [{k: v for k, v in g.items() if k != 'params'} for g in sd['param_groups']]
    return vals, groups


def same_everywhere(opt, what):
    mine = flat_state(opt)
    roots = hvd.broadcast_object(mine, root_rank=ROOT, name='chk.' + what)
    assert mine == roots, '%s: rank %d differs from root\n mine=%r\n root=%r' % (what, rank, mine[1], roots[1])


tested = []
for cls in classes():
    model = make_model(1)
    probe = cls(model.parameters(), lr=0.01) if 'lr' in inspect.signature(cls.__init__).parameters else cls(model.parameters())
    kw = perturb(probe.defaults, rank)
    opt = cls(model.parameters(), **kw)
    for step in range(1 + rank):                         # different step counts -> different state on every rank
        opt.zero_grad()
        model(torch.randn(4, 5, generator=torch.Generator().manual_seed(10 * rank + step))).pow(2).sum().backward()
        opt.step()
    hvd.broadcast_optimizer_state(opt, root_rank=ROOT)
    same_everywhere(opt, cls.__name__)
    # still trainable afterwards (load_state_dict kept dtypes / devices intact)
    opt.zero_grad()
    model(torch.randn(4, 5)).sum().backward()
    opt.step()
    tested.append(cls.__name__)

# state on the root only: the other ranks just constructed their optimizer
for cls in (torch.optim.Adam, torch.optim.SGD, torch.optim.RMSprop):
    model = make_model(2)
    kw = dict(lr=0.05 * (rank + 1), **({'momentum': 0.9} if cls is not torch.optim.Adam else {}))
    opt = cls(model.parameters(), **kw)
    if rank == ROOT:
        for _ in range(3):
            opt.zero_grad()
            model(torch.randn(4, 5)).sum().backward()
            opt.step()
    before = [p.detach().clone() for p in model.parameters()]
    hvd.broadcast_optimizer_state(opt, root_rank=ROOT)
    same_everywhere(opt, 'rootonly.' + cls.__name__)
    assert all(torch.equal(a, b) for a, b in zip(before, model.parameters())), 'materialising the state must not move the weights'
    assert all(p.grad is None for p in model.parameters()) or rank == ROOT

# wrapped optimizer
model = make_model(3)
opt = hvd.DistributedOptimizer(torch.optim.AdamW(model.parameters(), lr=0.001 * (rank + 1), betas=(0.9 - 0.01 * rank, 0.99)),
                               named_parameters=model.named_parameters())
opt.zero_grad()
model(torch.randn(4, 5)).sum().backward()
opt.step()
hvd.broadcast_optimizer_state(opt, root_rank=ROOT)
same_everywhere(opt, 'wrapped.AdamW')

try:
    hvd.broadcast_optimizer_state(torch.optim.LBFGS(make_model(4).parameters()), root_rank=0)
    raise SystemExit('LBFGS must be rejected')
except ValueError:
    pass

hvd.barrier()
if rank == 0:
    print('OPTIM STATE OK', len(tested), ' '.join(tested))
hvd.shutdown()
