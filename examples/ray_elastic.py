"""Elastic training on a Ray cluster (needs Ray; the reference's examples/ray/basic_ray_elastic.py).

    python examples/ray_elastic.py --min-workers 2 --max-workers 4

`RayExecutor(min_workers=..., max_workers=...)` discovers hosts from `ray.nodes()`, creates one actor per slot and re-plans
whenever nodes come or go; the training function is an ordinary `@hvd.elastic.run` function.  `callbacks=` receive whatever the
workers pass to `horovod_b200.ray.ray_logger.log`.  `--chaos` swaps the discovery for `TestDiscovery`, which removes and adds
hosts at random so the recovery path can be watched without touching the cluster.
"""
import argparse


def training_fn(epochs=5, steps=20):
    import torch
    import horovod_b200.torch as hvd
    from horovod_b200.ray import ray_logger
    hvd.init()
    torch.manual_seed(1234)
    model = torch.nn.Linear(4, 1)
    optimizer = hvd.DistributedOptimizer(torch.optim.SGD(model.parameters(), lr=0.05 * hvd.size()),
                                         named_parameters=model.named_parameters())
    state = hvd.elastic.TorchState(model, optimizer, epoch=0, batch=0)

    @hvd.elastic.run
    def train(state):
        for state.epoch in range(state.epoch, epochs):
            for state.batch in range(state.batch, steps):
                x = torch.randn(32, 4)
                loss = torch.nn.functional.mse_loss(model(x), x.sum(1, keepdim=True))
                optimizer.zero_grad()
                loss.backward()
                optimizer.step()
                if state.batch % 10 == 0:
                    state.commit()                    # a failure rolls back to here; new workers are synced from here
            ray_logger.log({'epoch': state.epoch, 'rank': hvd.rank(), 'size': hvd.size(), 'loss': float(loss)})
            state.batch = 0
            state.commit()
        return float(loss)
    return hvd.rank(), hvd.size(), train(state)


if __name__ == '__main__':
    p = argparse.ArgumentParser()
    p.add_argument('--min-workers', type=int, default=1)
    p.add_argument('--max-workers', type=int, default=None)
    p.add_argument('--use-gpu', action='store_true')
    p.add_argument('--chaos', action='store_true', help='TestDiscovery: hosts are removed and added at random')
    a = p.parse_args()
    import ray
    from horovod_b200.ray import RayExecutor
    ray.init(address='auto', ignore_reinit_error=True)
    settings = RayExecutor.create_settings(timeout_s=60)
    if a.chaos:
        from horovod_b200.ray.elastic_v2 import TestDiscovery
        settings.discovery = TestDiscovery(min_hosts=1, max_hosts=len(ray.nodes()), change_frequency_s=20, use_gpu=a.use_gpu)
    executor = RayExecutor(settings, min_workers=a.min_workers, max_workers=a.max_workers, use_gpu=a.use_gpu, reset_limit=20,
                           override_discovery=not a.chaos)
    executor.start()
    try:
        print(executor.run(training_fn, callbacks=[print]))
    finally:
        executor.shutdown()
