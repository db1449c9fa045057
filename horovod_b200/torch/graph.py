"""CUDA-graph training step: forward + backward captured once, replayed every step; gradients are reduced and the
(fused) optimizer applied right after the replay.

Why (B200-first; nothing like it exists in the reference): a ResNet-50 / BERT step in eager PyTorch issues 1–3 thousand
kernel launches, and on a B200 the GPU finishes them faster than one Python thread can issue them — the measured step
time is the host's launch time, identical at 1 and 8 GPUs.  Replaying the step as ONE graph launch removes that bound.
On several GPUs the gradient allreduce is PART of the graph: while the step is being captured, the hook of the last
gradient of every zero-copy bucket records that bucket's `hvd.captured_allreduce_` kernel (multimem / P2P in-place
reduction on peer-mapped memory, flag barrier inside the kernel) on a forked high-priority stream, so on replay the
reduction of bucket k overlaps the backward pass of the layers below it, and a whole data-parallel step — forward,
backward, every collective — is ONE launch with no negotiation round and no host work.  `optimizer.step()` after the
replay only applies the (fused) update.  If some trainable parameter is not in a registered bucket (sparse gradients,
`zero_copy=False`, gradient accumulation) the reductions stay outside the graph and `step()` issues them as before.

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
        self.comm_in_graph = False  # True: the gradient allreduces are kernel nodes of the captured graph
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
            self._agree_across_ranks()

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
        try:
            with torch.cuda.graph(self.graph):
                # multi-GPU: the gradient buckets' allreduce kernels become nodes of this graph (forked onto a side stream
                # as soon as a bucket's last gradient is recorded, joined before the capture ends)
                self.comm_in_graph = bool(getattr(opt, '_begin_graph_capture', lambda: False)())
                self.static_loss = self._fwd_bwd()
                if self.comm_in_graph:
                    opt._end_graph_capture()
        finally:
            if hasattr(opt, '_graph_capture'):
                opt._graph_capture = None
        torch.cuda.synchronize()
        self.captured = True

    def _agree_across_ranks(self):
        """A graph with the allreduce kernels inside only works if EVERY rank replays the same graph: if capture failed (or
        chose a different communication mode) anywhere, every rank drops to the eager step."""
        ps = getattr(self.optimizer, 'process_set', None)
        if ps is None or not ps.included() or (ps.size() or 1) <= 1:
            return
        from horovod_b200.torch import mpi_ops
        mine = torch.tensor([1.0 if self.captured else 0.0, 1.0 if self.comm_in_graph else 0.0], dtype=torch.float32)
        lo = mpi_ops.allreduce(mine, op=mpi_ops.Min, name='graphed_step.agree.min', process_set=ps)
        hi = mpi_ops.allreduce(mine, op=mpi_ops.Max, name='graphed_step.agree.max', process_set=ps)
        if self.captured and (lo[0] != hi[0] or lo[1] != hi[1]):
            self._abandon(RuntimeError('CUDA graph capture did not succeed identically on every rank'))

    def _abandon(self, exc):
        self.captured, self.graph, self.comm_in_graph = False, None, False
        self.optimizer._graph_mode = False
        if hasattr(self.optimizer, '_graph_comm_captured'):
            self.optimizer._graph_comm_captured = False
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
