"""Gradient compression for the TensorFlow front end (parity: horovod/tensorflow/compression.py:20-75; bf16 is new)."""
import tensorflow as tf


class Compressor:
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


def _cast_compressor(wire):
    class _C(Compressor):
        @staticmethod
        def compress(tensor):
            if tensor.dtype.is_floating and tensor.dtype != wire:
                return tf.cast(tensor, wire), tensor.dtype
            return tensor, None

        @staticmethod
        def decompress(tensor, ctx):
            return tensor if ctx is None else tf.cast(tensor, ctx)
    return _C


FP16Compressor = _cast_compressor(tf.float16)
BF16Compressor = _cast_compressor(tf.bfloat16)


class Compression:
    none = NoneCompressor
    fp16 = FP16Compressor
    bf16 = BF16Compressor
