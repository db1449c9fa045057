"""Gradient compression for the wire (reference horovod/torch/compression.py).

`Compression.fp16` / `Compression.bf16` cast floating tensors before the
collective and back afterwards. On the NVLink path the same effect is available
without the two extra elementwise kernels through HVD_WIRE_DTYPE=bf16|fp16
(cast fused into the allreduce kernel's pack/unpack ends)."""
import torch


class Compressor(object):
    """Interface for compressing and decompressing a given tensor."""

    @staticmethod
    def compress(tensor):
        raise NotImplementedError

    @staticmethod
    def decompress(tensor, ctx):
        raise NotImplementedError


class NoneCompressor(Compressor):
    @staticmethod
    def compress(tensor):
        return tensor, None

    @staticmethod
    def decompress(tensor, ctx):
        return tensor


class _CastCompressor(Compressor):
    wire_dtype = None

    @classmethod
    def compress(cls, tensor):
        if tensor.dtype.is_floating_point and tensor.dtype != cls.wire_dtype:
            return tensor.to(cls.wire_dtype), tensor.dtype
        return tensor, None

    @classmethod
    def decompress(cls, tensor, ctx):
        return tensor.to(ctx) if ctx is not None else tensor


class FP16Compressor(_CastCompressor):
    wire_dtype = torch.float16


class BF16Compressor(_CastCompressor):
    wire_dtype = torch.bfloat16


class Compression(object):
    """Optional gradient compression algorithm used during allreduce."""
    none = NoneCompressor
    fp16 = FP16Compressor
    bf16 = BF16Compressor
