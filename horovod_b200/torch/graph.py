"""CUDA-graph training step: forward + backward captured once, replayed every step; gradients are reduced and the
(fused) optimizer applied right after the replay.

Why (B200-first; nothing like it exists in the reference): a ResNet-50 / BERT step in eager PyTorch issues 1–3 thousand
kernel launches, and on a B200 the GPU finishes them faster than one Python thread can issue them — the measured step
time is the host's launch time, identical at 1 and 8 GPUs.  Replaying the step as ONE graph launch removes that bound.
The gradient hooks of `DistributedOptimizer` cannot fire from inside a replay, so in graph mode `optimizer.step()`
launches the allreduce of every (zero-copy) gradient bucket itself: ~100 MB of fp32 gradients take ≈0.3 ms on the NVLS
path of an 8×B200 NVSwitch box, so losing the overlap with backward costs less than the hooks' host time did.

    step = hvd.GraphedStep(lambda x, y: F.cross_entropy(model(x), y), optimizer, (x0, y0))
    for x, y in loader:
        loss = step(x, y)          # copies x, y into the static inputs, replays, allreduces, applies the update

Constraints (those of CUDA graphs): fixed input shapes/dtypes, no host synchronisation or data-dependent control flow
inside `step_fn`, `optimizer.zero_grad()` is part of the graph (do not call it yourself).  If capture fails the object
falls back to the eager step and says so in `.captured` / `.fallback_reason`.
"""
import warnings

import torch


class GraphedStep:
    """Training step whose forward + backward are one CUDA graph replay; `step(*inputs)` returns the (static) loss tensor.
    `.captured` tells whether the graph is in use, `.fallback_reason` why not."""

    def __init__(self, step_fn, optimizer, example_inputs, warmup_iters=3, enabled=True):
        self.step_fn, self.optimizer = step_fn, optimizer
        self.captured, self.fallback_reason = False, None
        self.graph, self.static_loss = None, None
        self.static_inputs = tuple(example_inputs)
        self.replays = 0
        if not enabled:
            self.fallback_reason = 'disabled'
        elif not (torch.cuda.is_available() and all(t.is_cuda for t in example_inputs)):
            self.fallback_reason = 'inputs are not CUDA tensors'
        elif not hasattr(optimizer, '_graph_mode'):
            self.fallback_reason = 'optimizer is not an hvd.DistributedOptimizer'
        else:
            try:
                self._capture(warmup_iters)
            except Exception as e:  # noqa: BLE001 - any capture failure must leave a working eager step behind
                self._abandon(e)

    # ------------------------------------------------------------------------------------------------------------------
    def _fwd_bwd(self):
        self.optimizer.zero_grad(set_to_none=False)
        loss = self.step_fn(*self.static_inputs)
        loss.backward()
        return loss

    def _capture(self, warmup_iters):
        opt = self.optimizer
        opt._graph_mode = True
        self.static_inputs = tuple(t.detach().clone(memory_format=torch.preserve_format) for t in self.static_inputs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup_iters)):  # cuDNN/cuBLAS autotuning, lazy allocations, gradient tensors
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        missing = [p for g in opt.param_groups for p in g['params'] if p.requires_grad and p.grad is None]
        for p in missing:  # parameters unused by step_fn still need a static gradient for the optimizer
            p.grad = torch.zeros_like(p, memory_format=torch.preserve_format)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._fwd_bwd()
        torch.cuda.synchronize()
        self.captured = True

    def _abandon(self, exc):
        self.captured, self.graph = False, None
        self.optimizer._graph_mode = False
        self.fallback_reason = '%s: %s' % (type(exc).__name__, str(exc).splitlines()[0] if str(exc) else '')
        try:
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
        warnings.warn('hvd.GraphedStep: CUDA graph capture failed (%s); running the eager step' % self.fallback_reason)

    # ------------------------------------------------------------------------------------------------------------------
    def __call__(self, *inputs):
        if not self.captured:
            self.optimizer.zero_grad()
            loss = self.step_fn(*inputs)
            loss.backward()
            self.optimizer.step()
            return loss
        for dst, src in zip(self.static_inputs, inputs):
            if dst is not src:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        self.optimizer.step()
        return self.static_loss
