"""Gradient compression for NDArrays (reference horovod/mxnet/compression.py: Compressor :18, NoneCompressor :31,
FP16Compressor :44, Compression :65)."""


class Compressor:
    """compress(tensor) -> (tensor on the wire, context); decompress(tensor, context) -> tensor in its original dtype."""

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


class FP16Compressor(Compressor):
    """Floating-point gradients travel as float16 and come back in their own dtype."""

    @staticmethod
    def compress(tensor):
        dt = str(tensor.dtype)
        if 'float' in dt and 'float16' not in dt:
            return tensor.astype('float16'), tensor.dtype
        return tensor, None

    @staticmethod
    def decompress(tensor, ctx):
        return tensor if ctx is None else tensor.astype(ctx)


class Compression:
    """Optional gradient compression algorithm used during allreduce."""
    none = NoneCompressor
    fp16 = FP16Compressor
