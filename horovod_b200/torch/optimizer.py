"""DistributedOptimizer: wraps any torch optimizer so that gradients are
allreduced (averaged) across ranks while the backward pass is still running.

API parity: horovod/torch/optimizer.py (DistributedOptimizer factory,
_DistributedOptimizer, _DistributedAdasumOptimizer).  Mechanism differences:
hooks use `register_post_accumulate_grad_hook` (fires right after AccumulateGrad
with no autograd-graph surgery), completion is event-chained (no host sync per
tensor), and `fused=True` replaces the wrapped SGD/Adam(W) math by one
multi-tensor sm_100a kernel (csrc/kernels/optim_kernels.cu).
"""
import os
import warnings
from contextlib import contextmanager

import torch

from horovod_b200.common.exceptions import HorovodInternalError
from horovod_b200.common.process_sets import global_process_set
from horovod_b200.common.util import split_list
from horovod_b200.torch import mpi_ops
from horovod_b200.torch.compression import Compression
from horovod_b200.torch.mpi_ops import (Adasum, Average, Sum, allreduce_async_, grouped_allreduce_async_, rank, size,
                                        synchronize)


class _DistributedOptimizer(torch.optim.Optimizer):
    def __init__(self, params, named_parameters, compression, backward_passes_per_step=1, op=Average,
                 gradient_predivide_factor=1.0, groups=None, sparse_as_dense=False, process_set=global_process_set,
                 fused=False, zero_copy=None, bucket_cap_mb=32, bucket_wire_dtype=None):
        super(self.__class__, self).__init__(params)
        self._compression = compression
        # opt-in: fp32 gradient buckets travel as bf16 / fp16 (a registered shadow bucket of half the bytes is what the
        # zero-copy kernel reduces; the switch accumulates in fp32).  Not validated on hardware yet (docs/roadmap.md B3).
        if bucket_wire_dtype is None:
            bucket_wire_dtype = {'bf16': torch.bfloat16, 'fp16': torch.float16}.get(os.environ.get('HVD_BUCKET_WIRE_DTYPE', ''))
        self._bucket_wire_dtype = bucket_wire_dtype

        if named_parameters is not None:
            named_parameters = list(named_parameters)
        else:
            named_parameters = [(f'allreduce.noname.{i}.{j}', v)
                                for i, param_group in enumerate(self.param_groups)
                                for j, v in enumerate(param_group['params'])]
        # make sure that named_parameters are tuples
        if any([not isinstance(p, tuple) for p in named_parameters]):
            raise ValueError('named_parameters should be a sequence of tuples (name, parameter), usually produced by '
                             'model.named_parameters().')
        dups = _find_duplicates([k for k, _ in named_parameters])
        if len(dups) > 0:
            raise ValueError('Parameter names in named_parameters must be unique. Found duplicates: %s' % ', '.join(dups))
        all_param_ids = {id(v) for param_group in self.param_groups for v in param_group['params']}
        named_param_ids = {id(v) for k, v in named_parameters}
        unnamed_param_ids = all_param_ids - named_param_ids
        if len(unnamed_param_ids):
            raise ValueError('named_parameters was specified, but one or more model parameters were not named. Python '
                             'object ids: %s' % ', '.join(str(id) for id in unnamed_param_ids))

        self._parameter_names = {v: k for k, v in sorted(named_parameters)}
        self.backward_passes_per_step = backward_passes_per_step
        self._allreduce_delay = {v: self.backward_passes_per_step for _, v in sorted(named_parameters)}
        self.op = op
        self.gradient_predivide_factor = gradient_predivide_factor
        self.sparse_as_dense = sparse_as_dense
        self.process_set = process_set
        self._fused = fused
        self._fused_steps = 0
        self._graph_mode = False  # set by hvd.GraphedStep: backward runs as a CUDA graph replay, hooks do not fire
        self._graph_capture = None      # while hvd.GraphedStep captures: {'comm': side stream, 'ctas': n, 'launched': set()}
        self._graph_comm_captured = False  # the captured graph contains the gradient allreduces (nothing left for step())

        self._handles = {}
        self._grad_accs = []
        self._requires_update = set()
        self._synchronized = False
        self._should_synchronize = True

        self._num_groups = None
        self._groups = None
        if groups is not None:
            if not (isinstance(groups, list) or groups > 0):
                raise ValueError('groups should be a non-negative integer or a list of list of torch.Tensor.')
            if isinstance(groups, list):
                grouped_parameter_ids = set()
                for l in groups:
                    for p in l:
                        if not isinstance(p, torch.Tensor):
                            raise ValueError('groups must consist of torch.Tensor.')
                        if id(p) in grouped_parameter_ids:
                            raise ValueError('A parameter can only appear once in groups.')
                        grouped_parameter_ids.add(id(p))
                self._groups = groups
            else:
                self._num_groups = groups
        self._p_to_group = {}
        self._group_counts = {}
        self._group_handles = {}
        self._buckets = []
        self._p_to_bucket = {}
        self._zero_copy = False
        if zero_copy is None:
            zero_copy = os.environ.get('HVD_ZERO_COPY', '1') != '0'
        if (zero_copy and groups is None and compression is Compression.none and op in (Average, Sum) and not sparse_as_dense
                and self.process_set.included() and mpi_ops.symm_available(self.process_set)):
            self._setup_buckets(int(bucket_cap_mb * 1024 * 1024))

        if self.process_set.included() and (size() > 1 or True):
            self._register_hooks()

    # -- reference API ----------------------------------------------------------
    def load_state_dict(self, *args, **kwargs):
        self._handles = {}
        self._synchronized = False
        self._should_synchronize = True
        for p in self._allreduce_delay:
            self._allreduce_delay[p] = self.backward_passes_per_step
        super(self.__class__, self).load_state_dict(*args, **kwargs)

    @staticmethod
    def find_duplicates(lst):
        return _find_duplicates(lst)

    def set_backward_passes_per_step(self, passes):
        self.backward_passes_per_step = passes
        for p in self._allreduce_delay:
            self._allreduce_delay[p] = self.backward_passes_per_step

    def _register_hooks(self):
        if self._groups is not None:
            p_list = []
            # identify how many gradients each group expects
            for i, group in enumerate(self._groups):
                for p in group:
                    self._p_to_group[p] = group
                    p_list.append(p)
                self._group_counts[id(group)] = 0
            for param_group in self.param_groups:
                for p in param_group['params']:
                    if p.requires_grad and p not in self._p_to_group:
                        self._p_to_group[p] = [p]
                        self._group_counts[id(self._p_to_group[p])] = 0
        elif self._num_groups:
            p_list = []
            for param_group in self.param_groups:
                for p in param_group['params']:
                    if p.requires_grad:
                        p_list.append(p)
            # rank-consistent grouping: order by name everywhere
            p_list = sorted(p_list, key=lambda p: self._parameter_names.get(p))
            self._groups = [list(g) for g in split_list(p_list, self._num_groups)]
            for group in self._groups:
                for p in group:
                    self._p_to_group[p] = group
                self._group_counts[id(group)] = 0

        for param_group in self.param_groups:
            for p in param_group['params']:
                if p.requires_grad:
                    self._requires_update.add(p)
                    self._grad_accs.append(p.register_post_accumulate_grad_hook(self._make_hook(p)))

    # ---- zero-copy gradient buckets (B200-first: the reference allreduces per-parameter tensors through a fusion buffer) ----
    def _setup_buckets(self, cap_bytes):
        """Lays the gradients of all dense CUDA parameters out in a few flat buffers of registered symmetric memory
        (p.grad becomes a view, like DDP's gradient_as_bucket_view). A bucket is allreduced IN PLACE as one tensor by the
        zero-copy NVLink kernel as soon as its last gradient is accumulated: no per-tensor negotiation, no pack/unpack."""
        params = [p for g in self.param_groups for p in g['params']
                  if p.requires_grad and p.is_cuda and mpi_ops._is_dense(p) and p.dtype in (torch.float32, torch.bfloat16, torch.float16)]
        params.reverse()  # gradients become ready roughly in reverse registration order
        by_dtype = {}
        for p in params:
            by_dtype.setdefault(p.dtype, []).append(p)
        try:
            for dtype, ps in by_dtype.items():
                itemsize = ps[0].element_size()
                cur, cur_bytes = [], 0
                layouts = []
                for p in ps:
                    nbytes = (p.numel() * itemsize + 127) // 128 * 128
                    if cur and cur_bytes + nbytes > cap_bytes:
                        layouts.append((cur, cur_bytes))
                        cur, cur_bytes = [], 0
                    cur.append((p, cur_bytes))
                    cur_bytes += nbytes
                if cur:
                    layouts.append((cur, cur_bytes))
                for members, total in layouts:
                    flat = mpi_ops.symm_empty(total // itemsize, dtype=dtype, device=members[0][0].device, process_set=self.process_set)
                    flat.zero_()
                    bucket = {'flat': flat, 'params': [m[0] for m in members], 'pending': len(members),
                              'name': 'bucket.%d' % len(self._buckets), 'handle': None, 'shadow': None, 'shadow_views': None,
                              'offsets': {m[0]: m[1] for m in members}, 'views': {}}
                    if self._bucket_wire_dtype is not None and dtype == torch.float32:
                        shadow = mpi_ops.symm_empty(total // itemsize, dtype=self._bucket_wire_dtype, device=members[0][0].device,
                                                    process_set=self.process_set)
                        shadow.zero_()
                        bucket['shadow'] = shadow
                        bucket['shadow_views'] = {p: shadow[off // itemsize: off // itemsize + p.numel()].as_strided(p.size(), p.stride())
                                                  for p, off in members}
                    for p, off in members:
                        view = flat[off // itemsize: off // itemsize + p.numel()].as_strided(p.size(), p.stride())
                        if p.grad is not None:
                            view.copy_(p.grad)
                        p.grad = view
                        bucket['views'][p] = view
                        self._p_to_bucket[p] = bucket
                    self._buckets.append(bucket)
            self._zero_copy = True
        except HorovodInternalError as e:
            warnings.warn('zero-copy gradient buckets unavailable (%s); using the fused pack/unpack path' % e)
            for b in self._buckets:
                for p in b['params']:
                    p.grad = None
            self._buckets, self._p_to_bucket, self._zero_copy = [], {}, False

    def _launch_bucket(self, bucket):
        if self.op == Average:
            prescale_factor, postscale_factor = 1.0 / self.gradient_predivide_factor, self.gradient_predivide_factor
        else:
            prescale_factor = postscale_factor = 1.0
        wire = bucket['flat']
        if bucket.get('shadow') is not None:
            wire = bucket['shadow']
            wire.copy_(bucket['flat'])  # one cast kernel fp32 -> wire dtype on the caller's stream
        bucket['handle'] = allreduce_async_(wire, name=bucket['name'], op=self.op, prescale_factor=prescale_factor,
                                            postscale_factor=postscale_factor, process_set=self.process_set)
        bucket['pending'] = len(bucket['params'])

    def _allreduce_grad_async(self, p):
        if p.grad is None:
            # gradient was not computed on this rank but the peers will reduce it: contribute zeros
            p.grad = torch.zeros_like(p.data)
        name = self._parameter_names.get(p)
        tensor = p.grad
        if tensor.is_sparse:
            if self.sparse_as_dense:
                tensor = tensor.to_dense()
                p.grad = tensor
            else:
                return mpi_ops.sparse_allreduce_async(tensor, name=name, op=self.op, process_set=self.process_set), None
        tensor_compressed, ctx = self._compression.compress(tensor)
        if self.op == Average:
            # split the averaging into pre- and post-division around the sum (reference optimizer.py:197-204)
            prescale_factor = 1.0 / self.gradient_predivide_factor
            postscale_factor = self.gradient_predivide_factor
        else:
            prescale_factor = 1.0
            postscale_factor = 1.0
        handle = allreduce_async_(tensor_compressed, name=name, op=self.op, prescale_factor=prescale_factor,
                                  postscale_factor=postscale_factor, process_set=self.process_set)
        return handle, ctx

    def _grouped_allreduce_grad_async(self, ps):
        name = self._parameter_names.get(ps[0])
        for p in ps:
            if p.grad is None:
                p.grad = torch.zeros_like(p.data)
        tensors_compressed, ctxs = zip(*[self._compression.compress(p.grad) for p in ps])
        if self.op == Average:
            prescale_factor = 1.0 / self.gradient_predivide_factor
            postscale_factor = self.gradient_predivide_factor
        else:
            prescale_factor = 1.0
            postscale_factor = 1.0
        handle = grouped_allreduce_async_(list(tensors_compressed), name=name, op=self.op, prescale_factor=prescale_factor,
                                          postscale_factor=postscale_factor, process_set=self.process_set)
        return handle, ctxs

    def _launch_group(self, group):
        handle, ctxs = self._grouped_allreduce_grad_async(group)
        self._group_handles[handle] = (group, ctxs)
        for gp in group:
            self._handles[gp] = (handle, None)
        self._group_counts[id(group)] = 0

    def _make_hook(self, p):
        def hook(*ignore):
            if self._graph_mode:
                # hooks only run during warm-up and capture; while capturing, the last gradient of a bucket records the
                # bucket's allreduce INTO the graph (on a forked side stream, so it overlaps the rest of backward)
                if self._graph_capture is not None:
                    self._graph_hook(p)
                return
            if p in self._handles and self._handles[p][0] is not None:
                if self._allreduce_delay[p] <= 0:
                    raise AssertionError(
                        "Gradients were computed more than backward_passes_per_step times before call to step(). "
                        "Increase backward_passes_per_step to accumulate gradients locally.")
            assert not p.grad.requires_grad
            assert self._allreduce_delay[p] > 0
            handle, ctx = None, None
            self._allreduce_delay[p] -= 1
            bucket = self._p_to_bucket.get(p)
            if bucket is not None:
                if p.grad.is_sparse or p.grad.data_ptr() != bucket['flat'].data_ptr() + bucket['offsets'][p]:
                    # model.zero_grad() (set_to_none=True is torch's default) or `p.grad = None` dropped the bucket view and
                    # autograd allocated a fresh gradient: move it into the registered bucket and re-point p.grad, so the
                    # reference idiom keeps working (one extra copy for that parameter; optimizer.zero_grad() avoids it)
                    view = bucket['views'][p]
                    view.copy_(p.grad.to_dense() if p.grad.is_sparse else p.grad)
                    p.grad = view
                self._handles[p] = ('bucket', None)
                if self._allreduce_delay[p] == 0:
                    bucket['pending'] -= 1
                    if bucket['pending'] == 0:
                        self._launch_bucket(bucket)
                return
            if self._allreduce_delay[p] == 0:
                if self._groups is not None:
                    group = self._p_to_group[p]
                    self._group_counts[id(group)] += 1
                    if self._group_counts[id(group)] == len(group):
                        self._launch_group(group)
                        return
                else:
                    handle, ctx = self._allreduce_grad_async(p)
            self._handles[p] = (handle, ctx)
        return hook

    # ---- gradient allreduce as CUDA-graph nodes (hvd.GraphedStep) --------------------------------------------------------
    def _graph_capturable(self):
        """The whole gradient reduction can live inside the captured step: every trainable parameter sits in a registered
        zero-copy bucket, one backward per step, a reduction the in-place kernel implements.  Depends on the model and the
        constructor arguments only, so every rank answers the same."""
        if not (self._zero_copy and self._buckets and self.backward_passes_per_step == 1 and self.op in (Average, Sum)):
            return False
        if not self.process_set.included() or self.process_set.size() <= 1:
            return False
        if os.environ.get('HVD_GRAPH_COMM', '1') == '0':
            return False
        return all(p in self._p_to_bucket for p in self._requires_update)

    def _begin_graph_capture(self):
        """Called by hvd.GraphedStep right after stream capture began (on the capturing stream)."""
        self._graph_comm_captured = False
        if not self._graph_capturable():
            return False
        lo_pri, hi_pri = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, 'priority_range') else (0, -1)
        self._graph_capture = {'comm': torch.cuda.Stream(priority=hi_pri), 'launched': set(),
                               # few CTAs while backward still needs the SMs (8 x B200, BERT-large: 16 CTAs 30.64 ms/step, 32: 30.73,
                               # 64: 31.82), all of them for the tail bucket(s) that nothing overlaps any more
                               # (without NVLS — teams of 2 — the P2P two-shot kernel needs more CTAs to keep up with backward)
                               'ctas': int(os.environ.get('HVD_GRAPH_COMM_CTAS', '16' if '+mc' in mpi_ops.gpu_backend_info() and self.process_set.size() >= 4 else '32')),
                               'tail_ctas': int(os.environ.get('HVD_GRAPH_COMM_TAIL_CTAS', '128')), 'tail': False,
                               'pending': {id(b): len(b['params']) for b in self._buckets}}
        return True

    def _graph_hook(self, p):
        bucket = self._p_to_bucket.get(p)
        if bucket is None:
            return
        cap = self._graph_capture
        cap['pending'][id(bucket)] -= 1
        if cap['pending'][id(bucket)] == 0:
            self._graph_launch_bucket(bucket)

    def _graph_launch_bucket(self, bucket):
        cap = self._graph_capture
        if self.op == Average:
            prescale_factor, postscale_factor = 1.0 / self.gradient_predivide_factor, self.gradient_predivide_factor
        else:
            prescale_factor = postscale_factor = 1.0
        comm = cap['comm']
        comm.wait_stream(torch.cuda.current_stream())  # fork: everything recorded so far (this bucket's gradients) precedes it
        with torch.cuda.stream(comm):
            wire = bucket['flat']
            if bucket.get('shadow') is not None:
                wire = bucket['shadow']
                wire.copy_(bucket['flat'])
            # the last bucket in registration order holds the first layers' gradients: backward is over when it is ready
            tail = cap['tail'] or bucket is self._buckets[-1]
            mpi_ops.captured_allreduce_(wire, op=self.op, prescale_factor=prescale_factor, postscale_factor=postscale_factor,
                                        process_set=self.process_set, max_ctas=cap['tail_ctas'] if tail else cap['ctas'])
        cap['launched'].add(id(bucket))

    def _end_graph_capture(self):
        """Before stream capture ends: buckets whose last hook never fired (parameters unused by the step) are reduced
        now, then the capturing stream joins the communication stream."""
        cap = self._graph_capture
        if cap is None:
            return
        cap['tail'] = True
        for bucket in self._buckets:
            if id(bucket) not in cap['launched']:
                self._graph_launch_bucket(bucket)
        torch.cuda.current_stream().wait_stream(cap['comm'])
        self._graph_capture = None
        self._graph_comm_captured = True

    def _adopt_missing_grads(self, bucket):
        """A bucket is launched from synchronize() because some hook never fired: parameters whose gradient is None (dropped
        by model.zero_grad() and not produced this step) contribute zeros and get their bucket view back."""
        base = bucket['flat'].data_ptr()
        for p in bucket['params']:
            view = bucket['views'][p]
            if p.grad is None:
                if p not in self._handles:
                    view.zero_()
                p.grad = view
            elif p.grad.is_sparse or p.grad.data_ptr() != base + bucket['offsets'][p]:
                view.copy_(p.grad.to_dense() if p.grad.is_sparse else p.grad)
                p.grad = view

    def _restore_shadows(self):
        """Reduced values of wire-dtype shadow buckets -> the fp32 gradients that p.grad points at."""
        for bucket in self._buckets:
            if bucket.get('shadow') is not None and bucket.get('shadow_fresh'):
                bucket['flat'].copy_(bucket['shadow'])
                bucket['shadow_fresh'] = False

    def _grad_for_update(self, p):
        """The tensor the fused optimizer kernel should read as p's gradient (the wire-dtype shadow when it is fresh)."""
        bucket = self._p_to_bucket.get(p)
        if bucket is not None and bucket.get('shadow') is not None and bucket.get('shadow_fresh'):
            return bucket['shadow_views'][p]
        return p.grad

    def synchronize(self, _defer_shadow_copy=False):
        """Waits for every outstanding gradient allreduce (enqueueing the ones whose hook never fired so that all
        ranks stay in lock-step) and writes the reduced gradients back."""
        if not self.process_set.included():
            self._synchronized = True
            return
        if self._graph_mode and self.process_set.size() == 1 and self.op in (Average, Sum):
            self._synchronized = True  # one rank: the reduction is the identity (prescale * postscale == 1)
            return
        if self._graph_mode and self._graph_comm_captured:
            # the replayed graph already reduced every bucket (captured_allreduce_ nodes): nothing to negotiate or launch
            for bucket in self._buckets:
                if bucket.get('shadow') is not None:
                    bucket['shadow_fresh'] = True
            if not _defer_shadow_copy:
                self._restore_shadows()
            self._synchronized = True
            return
        if self._zero_copy:
            for bucket in self._buckets:
                if bucket['handle'] is None:
                    # some gradient of this bucket was not produced on this rank (or accumulation is incomplete): reduce what
                    # is there so that all ranks stay in lock-step (missing gradients are zeros)
                    self._adopt_missing_grads(bucket)
                    self._launch_bucket(bucket)
            for bucket in self._buckets:
                synchronize(bucket['handle'])
                bucket['handle'] = None
                if bucket.get('shadow') is not None:
                    bucket['shadow_fresh'] = True
                for p in bucket['params']:
                    self._allreduce_delay[p] = self.backward_passes_per_step
                    self._handles.pop(p, None)
        if self._zero_copy and not _defer_shadow_copy:
            self._restore_shadows()  # callers of synchronize() (gradient clipping, ...) read p.grad
        pending = [p for p in self._requires_update if p not in self._handles and p not in self._p_to_bucket]
        pending += [p for p, (h, _) in self._handles.items() if h is None]
        if self._groups is not None:
            launched = set()
            for p in pending:
                group = self._p_to_group[p]
                if id(group) not in launched:
                    launched.add(id(group))
                    self._launch_group(group)
        else:
            for p in pending:
                self._handles[p] = self._allreduce_grad_async(p)

        waited = set()
        for p, (handle, ctx) in self._handles.items():
            if handle in self._group_handles:
                if handle in waited:
                    continue
                waited.add(handle)
                group, ctxs = self._group_handles[handle]
                outputs = synchronize(handle)
                for gp, out, c in zip(group, outputs, ctxs):
                    self._allreduce_delay[gp] = self.backward_passes_per_step
                    if out is not gp.grad:
                        gp.grad.set_(self._compression.decompress(out, c))
                continue
            output = synchronize(handle)
            self._allreduce_delay[p] = self.backward_passes_per_step
            if p.grad.is_sparse:
                aggregated = self._compression.decompress(output, ctx)
                if not aggregated.is_sparse:
                    aggregated = aggregated.to_sparse()
                p.grad = aggregated
            elif output is not None and output is not p.grad:
                p.grad.set_(self._compression.decompress(output, ctx))
        self._handles.clear()
        self._group_handles.clear()
        self._synchronized = True

    @contextmanager
    def skip_synchronize(self):
        """Use after an explicit optimizer.synchronize() (e.g. gradient clipping) so step() does not sync again."""
        self._should_synchronize = False
        try:
            yield
        finally:
            self._should_synchronize = True

    def _hvd_super_step(self, closure=None):
        return super(self.__class__, self).step(closure)

    def step(self, closure=None):
        if self._should_synchronize:
            if self._synchronized:
                warnings.warn("optimizer.step() called without optimizer.skip_synchronize() context after "
                              "optimizer.synchronize(). This can cause training slowdown. You may want to consider "
                              "using optimizer.skip_synchronize() context if you use optimizer.synchronize() in your "
                              "code.")
            self.synchronize(_defer_shadow_copy=self._fused and closure is None)
        self._synchronized = False
        if self._fused and closure is None and _fused_step(self):
            for bucket in self._buckets:
                bucket['shadow_fresh'] = False
            return None
        self._restore_shadows()
        return super(self.__class__, self).step(closure)

    def zero_grad(self, *args, **kwargs):
        if self._handles:
            raise AssertionError("optimizer.zero_grad() was called after loss.backward() but before optimizer.step() or "
                                 "optimizer.synchronize(). This is prohibited as it can cause a race condition.")
        if self._zero_copy:
            # bucketed gradients must stay views of the registered buffers: zero them in place (one memset per bucket)
            for bucket in self._buckets:
                bucket['flat'].zero_()
            for group in self.param_groups:
                for p in group['params']:
                    if p not in self._p_to_bucket and p.grad is not None:
                        if self._graph_mode:
                            p.grad.zero_()  # addresses baked into the captured graph must stay valid
                        else:
                            p.grad = None
            return None
        if self._graph_mode:
            kwargs['set_to_none'] = False
            args = ()
        return super(self.__class__, self).zero_grad(*args, **kwargs)


def _find_duplicates(lst):
    seen = set()
    dups = set()
    for el in lst:
        if el in seen:
            dups.add(el)
        seen.add(el)
    return dups


def _fused_plan(opt):
    """Pure eligibility pass of the fused optimizer step: returns the launch plan [(group, params, grads)] or None.
    Nothing is mutated here, so a `None` leaves the wrapped optimizer's own step() as a clean fallback."""
    is_sgd = isinstance(opt, torch.optim.SGD)
    is_adam = isinstance(opt, (torch.optim.Adam, torch.optim.AdamW))
    if not (is_sgd or is_adam):
        return None
    grad_of = getattr(opt, '_grad_for_update', lambda q: q.grad)
    plan = []
    for group in opt.param_groups:
        params = [p for p in group['params'] if p.grad is not None]
        if not params:
            continue
        if group.get('maximize', False) or (is_adam and group.get('amsgrad', False)):
            return None
        if not all(p.is_cuda and not p.grad.is_sparse and mpi_ops._is_dense(p) and p.grad.stride() == p.stride() for p in params):
            return None
        by_dtype = {}
        for p in params:
            by_dtype.setdefault((p.dtype, grad_of(p).dtype), []).append(p)
        if any(pd == torch.float32 and gd == torch.float16 for pd, gd in by_dtype):
            return None  # no fp32-parameter / fp16-gradient kernel
        for ps in by_dtype.values():
            if is_sgd and group.get('momentum', 0.0) != 0.0:
                have = [opt.state[p].get('momentum_buffer') is not None for p in ps]
                if any(have) and not all(have):
                    return None  # partially initialised momentum (parameters added later): let torch handle it
            plan.append((group, ps, [grad_of(p) for p in ps]))
    return plan


def _fused_step(opt):
    """One multi-tensor kernel for the whole model when the wrapped optimizer is SGD or Adam/AdamW on CUDA."""
    native = mpi_ops._native()
    plan = _fused_plan(opt)
    if plan is None:
        return False
    is_sgd = isinstance(opt, torch.optim.SGD)
    with torch.no_grad():
        for group, ps, grads in plan:
            if is_sgd:
                mom = group.get('momentum', 0.0)
                bufs, first = [], False
                if mom != 0.0:
                    for p in ps:
                        st = opt.state[p]
                        if st.get('momentum_buffer') is None:
                            st['momentum_buffer'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                            first = True
                        bufs.append(st['momentum_buffer'])
                native.fused_sgd_step(ps, grads, bufs, float(group['lr']), float(mom), float(group.get('dampening', 0.0)),
                                      float(group.get('weight_decay', 0.0)), bool(group.get('nesterov', False)), 1.0, first)
            else:
                exp_avg, exp_avg_sq, steps = [], [], []
                for p in ps:
                    st = opt.state[p]
                    if len(st) == 0 or 'exp_avg' not in st:
                        st['step'] = torch.tensor(0.0)
                        st['exp_avg'] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.preserve_format)
                        st['exp_avg_sq'] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.preserve_format)
                    if not torch.is_tensor(st['step']):
                        st['step'] = torch.tensor(float(st['step']))
                    steps.append(st['step'])
                    exp_avg.append(st['exp_avg'])
                    exp_avg_sq.append(st['exp_avg_sq'])
                # one host-side call for the whole group instead of a tensor op per parameter (BERT-large: ~400 of them)
                cpu_steps = [t for t in steps if not t.is_cuda]
                if cpu_steps:
                    torch._foreach_add_(cpu_steps, 1)
                for t in steps:
                    if t.is_cuda:
                        t.add_(1)
                step = int(steps[0].item())
                b1, b2 = group['betas']
                adamw = isinstance(opt, torch.optim.AdamW) or bool(group.get('decoupled_weight_decay', False))
                native.fused_adam_step(ps, grads, exp_avg, exp_avg_sq, float(group['lr']), float(b1), float(b2),
                                       float(group['eps']), float(group.get('weight_decay', 0.0)), step, 1.0, adamw)
    return True


class _DistributedAdasumOptimizer(torch.optim.Optimizer):
    """Adasum works on parameter *deltas*: each parameter's hook runs the wrapped optimizer on that parameter only,
    allreduces delta = p_new - p_start with op=Adasum and step() applies start + adasum(delta).
    (reference horovod/torch/optimizer.py:345-513)"""

    def __init__(self, params, named_parameters, compression, backward_passes_per_step=1):
        super(self.__class__, self).__init__(params)
        self._compression = compression
        if named_parameters is not None:
            named_parameters = list(named_parameters)
        else:
            named_parameters = [(f'allreduce.noname.{i}.{j}', v)
                                for i, param_group in enumerate(self.param_groups)
                                for j, v in enumerate(param_group['params'])]
        all_param_ids = {id(v) for param_group in self.param_groups for v in param_group['params']}
        named_param_ids = {id(v) for k, v in named_parameters}
        unnamed_param_ids = all_param_ids - named_param_ids
        if len(unnamed_param_ids):
            raise ValueError('named_parameters was specified, but one or more model parameters were not named. Python '
                             'object ids: %s' % ', '.join(str(id) for id in unnamed_param_ids))
        self._parameter_names = {v: k for k, v in sorted(named_parameters)}
        self.backward_passes_per_step = backward_passes_per_step
        self._allreduce_delay = {v: self.backward_passes_per_step for _, v in sorted(named_parameters)}
        self._handles = {}
        self._grad_accs = []
        self._requires_update = set()
        self._synchronized = False
        self._should_synchronize = True
        self._starting_models = {p: torch.zeros_like(p, requires_grad=False) for _, p in named_parameters}
        self._graph_mode = False  # hvd.GraphedStep: hooks are silent, step() runs the whole-model variant below
        self._register_hooks()

    def set_backward_passes_per_step(self, passes):
        self.backward_passes_per_step = passes
        for p in self._allreduce_delay:
            self._allreduce_delay[p] = self.backward_passes_per_step

    def _register_hooks(self):
        for param_group in self.param_groups:
            for p in param_group['params']:
                if p.requires_grad:
                    self._requires_update.add(p)
                    self._grad_accs.append(p.register_post_accumulate_grad_hook(self._make_hook(p)))

    def _allreduce_grad_async(self, p):
        # run the wrapped optimizer on this parameter alone
        name = self._parameter_names.get(p)
        if p.grad is None:
            p.grad = torch.zeros_like(p.data)
        stashed_params = []
        for group in self.param_groups:
            stashed_params.append(group['params'])
            if any(p is v for v in group['params']):
                group['params'] = [p]
            else:
                group['params'] = []
        start = self._starting_models[p]
        start.data.copy_(p)
        super(self.__class__, self).step()
        p.data.sub_(start)  # p now holds the local delta
        tensor_compressed, ctx = self._compression.compress(p.data)
        handle = allreduce_async_(tensor_compressed, name=name, op=Adasum)
        for stashed, group in zip(stashed_params, self.param_groups):
            group['params'] = stashed
        return handle, ctx

    def _make_hook(self, p):
        def hook(*ignore):
            if self._graph_mode:
                return
            if p in self._handles and self._handles[p][0] is not None:
                if self._allreduce_delay[p] <= 0:
                    raise AssertionError(
                        "Gradients were computed more than backward_passes_per_step times before call to step(). "
                        "Increase backward_passes_per_step to accumulate gradients locally.")
            assert not p.grad.requires_grad
            assert self._allreduce_delay[p] > 0
            handle, ctx = None, None
            self._allreduce_delay[p] -= 1
            if self._allreduce_delay[p] == 0:
                handle, ctx = self._allreduce_grad_async(p)
            self._handles[p] = (handle, ctx)
        return hook

    def _hvd_super_step(self, closure=None):
        return super(self.__class__, self).step(closure)

    def _whole_model_step(self):
        """Same math as the per-parameter hooks, issued once for the whole model (graph mode: backward was a CUDA graph
        replay, so there is nothing to overlap with): stash, ONE wrapped-optimizer step over all parameters, delta =
        new - start, Adasum-allreduce every delta (the engine fuses them up to the fusion threshold), start + result."""
        ps = sorted(self._requires_update, key=lambda p: self._parameter_names.get(p))
        for p in ps:
            if p.grad is None:
                p.grad = torch.zeros_like(p.data)
        starts = [self._starting_models[p] for p in ps]
        datas = [p.data for p in ps]
        with torch.no_grad():
            torch._foreach_copy_(starts, datas)
            super(self.__class__, self).step()
            torch._foreach_sub_(datas, starts)
            handles = []
            for p in ps:
                comp, ctx = self._compression.compress(p.data)
                handles.append((allreduce_async_(comp, name=self._parameter_names.get(p), op=Adasum), ctx, p))
            for h, ctx, p in handles:
                delta = self._compression.decompress(synchronize(h), ctx)
                if delta is not p.data:
                    p.data.copy_(delta)
            torch._foreach_add_(starts, datas)
            torch._foreach_copy_(datas, starts)

    def synchronize(self):
        pass

    @contextmanager
    def skip_synchronize(self):
        raise AssertionError("Skipping synchronization is not supported when using Adasum optimizer.")

    def step(self, closure=None):
        loss = None
        if closure is not None:
            loss = closure()
        if self._graph_mode:
            self._whole_model_step()
            return loss
        missing_p = self._requires_update - set(self._handles.keys())
        for p in missing_p:
            self._allreduce_delay[p] = 0
            handle, ctx = self._allreduce_grad_async(p)
            self._handles[p] = (handle, ctx)
        for p, (handle, ctx) in self._handles.items():
            if handle is None:
                handle, ctx = self._allreduce_grad_async(p)
                self._handles[p] = (handle, ctx)
        for p, (handle, ctx) in self._handles.items():
            delta = synchronize(handle)
            delta = self._compression.decompress(delta, ctx)
            start = self._starting_models[p]
            start.data.add_(delta.data)
            p.data.copy_(start)
            self._allreduce_delay[p] = self.backward_passes_per_step
        self._handles.clear()
        return loss

    def zero_grad(self, *args, **kwargs):
        if self._handles:
            raise AssertionError("optimizer.zero_grad() was called after loss.backward() but before optimizer.step() or "
                                 "optimizer.synchronize(). This is prohibited as it can cause a race condition.")
        return super(self.__class__, self).zero_grad(*args, **kwargs)


def DistributedOptimizer(optimizer, named_parameters=None, compression=Compression.none, backward_passes_per_step=1,
                         op=Average, gradient_predivide_factor=1.0, num_groups=0, groups=None, sparse_as_dense=False,
                         process_set=global_process_set, fused=False, zero_copy=None, bucket_cap_mb=32, bucket_wire_dtype=None):
    """Wraps `optimizer` so gradients are combined across ranks before the parameter update.

    Arguments follow the reference (horovod/torch/optimizer.py:516-608). `fused=True` (new) runs the SGD /
    Adam(W) update as one multi-tensor CUDA kernel."""
    if gradient_predivide_factor != 1.0:
        if rocm_built_safe():
            raise ValueError('gradient_predivide_factor not supported yet with ROCm')
        if op != Average:
            raise ValueError('gradient_predivide_factor not supported with op != Average')
    if num_groups != 0:
        warnings.warn('Parameter `num_groups` has been replaced by `groups` and will be removed', DeprecationWarning)
        if groups is None:
            groups = num_groups
    if op != Adasum or size() == 1:
        cls = type(optimizer.__class__.__name__, (optimizer.__class__,), dict(_DistributedOptimizer.__dict__))
        return cls(optimizer.param_groups, named_parameters, compression, backward_passes_per_step, op,
                   gradient_predivide_factor, groups, sparse_as_dense, process_set, fused, zero_copy, bucket_cap_mb, bucket_wire_dtype)
    if process_set != global_process_set:
        raise NotImplementedError("Adasum does not support non-global process sets yet.")
    cls = type(optimizer.__class__.__name__, (optimizer.__class__,), dict(_DistributedAdasumOptimizer.__dict__))
    return cls(optimizer.param_groups, named_parameters, compression, backward_passes_per_step)


def rocm_built_safe():
    try:
        return mpi_ops.rocm_built()
    except Exception:
        return False
