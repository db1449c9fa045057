"""`import horovod_b200.tensorflow.keras as hvd` (parity: horovod/tensorflow/keras/__init__.py:49-290)."""
import tensorflow as tf

import horovod_b200._keras as _impl
from horovod_b200.tensorflow import (  # noqa: F401
    init, shutdown, is_initialized, start_timeline, stop_timeline, size, local_size, cross_size, rank, local_rank,
    cross_rank, is_homogeneous, mpi_threads_supported, mpi_enabled, mpi_built, gloo_enabled, gloo_built, nccl_built,
    ddl_built, ccl_built, cuda_built, rocm_built, Average, Sum, Adasum, Min, Max, Product, global_process_set, ProcessSet,
    add_process_set, remove_process_set, Compression, broadcast_variables, broadcast_object, allgather_object,
    SyncBatchNormalization, PartialDistributedGradientTape, barrier)
from horovod_b200.tensorflow.keras import callbacks, elastic  # noqa: F401


def DistributedOptimizer(optimizer, name=None, device_dense='', device_sparse='', compression=Compression.none,
                         sparse_as_dense=False, gradient_predivide_factor=1.0, op=Average, backward_passes_per_step=1,
                         average_aggregated_gradients=False, num_groups=0, groups=None, process_set=global_process_set,
                         scale_local_gradients=True):
    if gradient_predivide_factor != 1.0 and op != Average:
        raise ValueError('gradient_predivide_factor not supported with op != Average')
    if op == Adasum and average_aggregated_gradients:
        raise ValueError('Adasum does not support average_aggregated_gradients == True')
    if num_groups != 0 and groups is None:
        groups = num_groups
    if groups is not None and not (isinstance(groups, list) or groups > 0):
        raise ValueError('groups should be a non-negative integer or a list of list of tf.Variable.')
    return _impl.create_distributed_optimizer(tf.keras, optimizer, name, device_dense, device_sparse, compression, sparse_as_dense,
                                              gradient_predivide_factor, op, backward_passes_per_step,
                                              average_aggregated_gradients, groups, process_set, scale_local_gradients)


def PartialDistributedOptimizer(optimizer, name=None, device_dense='', device_sparse='', compression=Compression.none,
                                sparse_as_dense=False, gradient_predivide_factor=1.0, op=Average, backward_passes_per_step=1,
                                average_aggregated_gradients=False, groups=None, process_set=global_process_set,
                                local_layers=None, scale_local_gradients=True):
    """DistributedOptimizer whose gradients for the variables of `local_layers` stay local (they are only scaled by
    1/size when `scale_local_gradients`): for model-parallel layers such as per-rank embedding shards (reference
    tensorflow/keras/__init__.py:173-230)."""
    if local_layers is None:
        local_layers = []
    elif not isinstance(local_layers, (list, tuple)):
        local_layers = [local_layers]
    if not all(hasattr(layer, 'trainable_weights') for layer in local_layers):
        raise ValueError('All local layers must be of tf.keras.layers.Layer type.')
    opt = DistributedOptimizer(optimizer, name=name, device_dense=device_dense, device_sparse=device_sparse, compression=compression,
                               sparse_as_dense=sparse_as_dense, gradient_predivide_factor=gradient_predivide_factor, op=op,
                               backward_passes_per_step=backward_passes_per_step,
                               average_aggregated_gradients=average_aggregated_gradients, groups=groups, process_set=process_set,
                               scale_local_gradients=scale_local_gradients)
    for layer in local_layers:
        for var in layer.trainable_weights:
            opt.register_local_var(var)
    return opt


def broadcast_global_variables(root_rank):
    return _impl.broadcast_global_variables(tf.keras.backend, root_rank)


def allreduce(value, name=None, average=None, prescale_factor=1.0, postscale_factor=1.0, op=None, compression=Compression.none):
    return _impl.allreduce(tf.keras.backend, value, name, average, prescale_factor, postscale_factor, op, compression)


def allgather(value, name=None):
    return _impl.allgather(tf.keras.backend, value, name)


def broadcast(value, root_rank, name=None):
    return _impl.broadcast(tf.keras.backend, value, root_rank, name)


def reducescatter(value, name=None, op=Average):
    return _impl.reducescatter(tf.keras.backend, value, name, op)


def load_model(filepath, custom_optimizers=None, custom_objects=None, compression=Compression.none, legacy_opts=False):
    def wrap(cls):
        return lambda **kwargs: DistributedOptimizer(cls(**kwargs), compression=compression)
    return _impl.load_model(tf.keras, wrap, filepath, custom_optimizers, custom_objects, legacy_opts)
