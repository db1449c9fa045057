"""RayExecutor: run an hvd job on a Ray cluster (or, without Ray, on local processes through the same job core).

    python examples/ray_executor.py --num-workers 2
    python examples/ray_executor.py --num-hosts 2 --num-workers-per-host 8 --use-gpu     # on a Ray cluster of 8-GPU nodes
"""
import argparse


def train(steps):
    import torch
    import horovod_b200.torch as hvd
    hvd.init()
    torch.manual_seed(0)
    model = torch.nn.Linear(4, 1)
    opt = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.1), named_parameters=model.named_parameters())
    hvd.broadcast_parameters(model.state_dict(), root_rank=0)
    x = torch.randn(64, 4, generator=torch.Generator().manual_seed(hvd.rank()))
    y = x.sum(1, keepdim=True)
    for _ in range(steps):
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(x), y)
        loss.backward()
        opt.step()
    result = (hvd.rank(), hvd.size(), round(loss.item(), 4), model.weight.detach().flatten().tolist())
    hvd.shutdown()
    return result


if __name__ == '__main__':
    p = argparse.ArgumentParser()
    p.add_argument('--num-workers', type=int, default=None)
    p.add_argument('--num-hosts', type=int, default=None)
    p.add_argument('--num-workers-per-host', type=int, default=1)
    p.add_argument('--use-gpu', action='store_true')
    p.add_argument('--steps', type=int, default=50)
    a = p.parse_args()
    from horovod_b200.ray import RayExecutor
    backend = None
    try:
        import ray
        ray.init(ignore_reinit_error=True)
    except ImportError:
        from horovod_b200.runner.cluster_job import LocalProcessBackend
        backend = LocalProcessBackend()
        print('Ray is not installed: running the same executor on local processes')
    if a.num_workers is None and a.num_hosts is None:
        a.num_workers = 2
    ex = RayExecutor(RayExecutor.create_settings(timeout_s=60), num_workers=a.num_workers, num_hosts=a.num_hosts,
                     num_workers_per_host=a.num_workers_per_host, use_gpu=a.use_gpu, backend=backend)
    ex.start()
    try:
        results = ex.run(train, args=[a.steps])
    finally:
        ex.shutdown()
    for r in results:
        print('rank %d of %d: loss %.4f' % r[:3])
    same = all(r[3] == results[0][3] for r in results)
    print('RAY EXAMPLE OK' if same and results[0][2] < 0.05 else 'ranks disagree or did not converge')
