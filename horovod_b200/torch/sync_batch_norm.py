"""SyncBatchNorm: batch statistics computed over the whole (global) batch.

Forward allgathers per-rank (count, mean, invstd) and combines them with
torch.batch_norm_gather_stats_with_counts; backward allreduces sum_dy and
sum_dy_xmu.  API parity: horovod/torch/sync_batch_norm.py."""
import torch
import torch.nn.functional as F
from torch.autograd.function import Function
from torch.nn.modules.batchnorm import _BatchNorm

from horovod_b200.torch.mpi_ops import Sum, allgather_async, allreduce_async, size, synchronize


class SyncBatchNorm(_BatchNorm):
    """Applies synchronous batch normalization: identical to torch.nn.BatchNormNd in eval mode and at size() == 1."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__(num_features, eps, momentum, affine, track_running_stats)

    def _check_input_dim(self, input):
        if input.dim() < 2:
            raise ValueError('expected at least 2D input (got {}D input)'.format(input.dim()))

    def _run_bn(self, input, factor):
        return F.batch_norm(input, self.running_mean, self.running_var, self.weight, self.bias,
                            self.training or not self.track_running_stats, factor, self.eps)

    @torch.jit.unused
    def _maybe_run_sync_bn(self, input, factor):
        if size() == 1:
            return self._run_bn(input, factor)
        return _SyncBatchNorm.apply(input, self.weight, self.bias, self.running_mean, self.running_var, self.eps, factor)

    def forward(self, input):
        self._check_input_dim(input)
        # exponential averaging factor as in torch.nn.modules.batchnorm._BatchNorm.forward: momentum=None means a cumulative
        # moving average over the batches seen so far
        factor = 0.0 if self.momentum is None else self.momentum
        if self.training and self.track_running_stats:
            self.num_batches_tracked = self.num_batches_tracked + 1
            if self.momentum is None:
                factor = 1.0 / float(self.num_batches_tracked)
        if not self.training and self.track_running_stats:
            return self._run_bn(input, factor)
        if not input.is_cuda and size() > 1:
            return _SyncBatchNormCPU.apply(input, self.weight, self.bias, self.running_mean, self.running_var, self.eps, factor)
        return self._maybe_run_sync_bn(input, factor)


class _SyncBatchNorm(Function):
    @staticmethod
    def forward(self, input, weight, bias, running_mean, running_var, eps, momentum):
        input = input.contiguous()
        size_ = input.numel() // input.size(1)
        count = torch.tensor([size_], device=input.device, dtype=torch.float32)
        mean, invstd = torch.batch_norm_stats(input, eps)
        count_handle = allgather_async(count.unsqueeze(0), name='sync_batch_norm.count')
        mean_handle = allgather_async(mean.unsqueeze(0), name='sync_batch_norm.mean')
        invstd_handle = allgather_async(invstd.unsqueeze(0), name='sync_batch_norm.invstd')
        count_all = synchronize(count_handle)
        mean_all = synchronize(mean_handle)
        invstd_all = synchronize(invstd_handle)
        counts = count_all.view(-1).float()
        mean, invstd = torch.batch_norm_gather_stats_with_counts(input, mean_all, invstd_all, running_mean, running_var,
                                                                 momentum, eps, counts)
        self.save_for_backward(input, weight, mean, invstd, count_all)
        return torch.batch_norm_elemt(input, weight, bias, mean, invstd, eps)

    @staticmethod
    def backward(self, grad_output):
        grad_output = grad_output.contiguous()
        saved_input, weight, mean, invstd, count_all = self.saved_tensors
        need_input_grad, need_weight_grad, need_bias_grad = self.needs_input_grad[0:3]
        sum_dy, sum_dy_xmu, grad_weight, grad_bias = torch.batch_norm_backward_reduce(
            grad_output, saved_input, mean, invstd, weight, need_input_grad, need_weight_grad, need_bias_grad)
        if need_input_grad:
            sum_dy_handle = allreduce_async(sum_dy, op=Sum, name='sync_batch_norm.sum_dy')
            sum_dy_xmu_handle = allreduce_async(sum_dy_xmu, op=Sum, name='sync_batch_norm.sum_dy_xmu')
            sum_dy = synchronize(sum_dy_handle)
            sum_dy_xmu = synchronize(sum_dy_xmu_handle)
            counts = count_all.view(-1).to(torch.int32)
            grad_input = torch.batch_norm_backward_elemt(grad_output, saved_input, mean, invstd, weight, sum_dy,
                                                         sum_dy_xmu, counts)
        else:
            grad_input = None
        if weight is None or not need_weight_grad:
            grad_weight = None
        if weight is None or not need_bias_grad:
            grad_bias = None
        return grad_input, grad_weight, grad_bias, None, None, None, None, None, None


class _SyncBatchNormCPU(Function):
    """Host fallback (the ATen batch_norm_stats kernels are CUDA-only): same math with plain tensor ops."""

    @staticmethod
    def forward(ctx, input, weight, bias, running_mean, running_var, eps, momentum):
        dims = [0] + list(range(2, input.dim()))
        n_local = input.numel() // input.size(1)
        local_sum = input.sum(dims)
        local_sqsum = (input * input).sum(dims)
        stats = torch.cat([local_sum, local_sqsum, torch.tensor([float(n_local)], dtype=input.dtype)])
        stats = synchronize(allreduce_async(stats, op=Sum, name='sync_batch_norm.cpu_stats'))
        c = input.size(1)
        n = stats[-1]
        mean = stats[:c] / n
        var = stats[c:2 * c] / n - mean * mean
        invstd = torch.rsqrt(var + eps)
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * mean)
            unbiased = var * n / (n - 1) if float(n) > 1 else var  # a single sample has no unbiased estimate
            running_var.mul_(1 - momentum).add_(momentum * unbiased)
        shape = [1, c] + [1] * (input.dim() - 2)
        xhat = (input - mean.view(shape)) * invstd.view(shape)
        ctx.save_for_backward(xhat, weight, invstd, n)
        out = xhat
        if weight is not None:
            out = out * weight.view(shape)
        if bias is not None:
            out = out + bias.view(shape)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        xhat, weight, invstd, n = ctx.saved_tensors
        c = xhat.size(1)
        dims = [0] + list(range(2, xhat.dim()))
        shape = [1, c] + [1] * (xhat.dim() - 2)
        grad_bias = grad_output.sum(dims)
        grad_weight = (grad_output * xhat).sum(dims)
        both = synchronize(allreduce_async(torch.cat([grad_bias, grad_weight]), op=Sum, name='sync_batch_norm.cpu_bwd'))
        sum_dy, sum_dy_xhat = both[:c], both[c:]
        g = grad_output * (weight.view(shape) if weight is not None else 1.0)
        w = weight if weight is not None else torch.ones(c, dtype=xhat.dtype)
        grad_input = (g - (w * sum_dy / n).view(shape) - xhat * (w * sum_dy_xhat / n).view(shape)) * invstd.view(shape)
        return grad_input, grad_weight if weight is not None else None, grad_bias, None, None, None, None
