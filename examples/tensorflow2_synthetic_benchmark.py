"""TensorFlow 2 custom training loop with `hvd.DistributedGradientTape` (synthetic data).

    hvdrun -np 8 python examples/tensorflow2_synthetic_benchmark.py --model ResNet50 --batch-size 64

Needs TensorFlow >= 2.4 (not part of this repository's build image, where the TensorFlow front end is exercised against a
numpy stand-in: tests/fakes/tensorflow).  GPU tensors reach the NVLink kernels through DLPack without a copy.
"""
import argparse
import time

import tensorflow as tf

import horovod_b200.tensorflow as hvd


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--model', default='ResNet50')
    p.add_argument('--batch-size', type=int, default=32)
    p.add_argument('--num-warmup-batches', type=int, default=5)
    p.add_argument('--num-batches-per-iter', type=int, default=10)
    p.add_argument('--num-iters', type=int, default=5)
    p.add_argument('--fp16-allreduce', action='store_true')
    a = p.parse_args()

    hvd.init()
    gpus = tf.config.list_physical_devices('GPU')
    if gpus:
        tf.config.set_visible_devices(gpus[hvd.local_rank()], 'GPU')
        tf.config.experimental.set_memory_growth(gpus[hvd.local_rank()], True)

    model = getattr(tf.keras.applications, a.model)(weights=None)
    opt = tf.keras.optimizers.SGD(0.01 * hvd.size())
    images = tf.random.uniform([a.batch_size, 224, 224, 3])
    labels = tf.random.uniform([a.batch_size], maxval=1000, dtype=tf.int64)
    loss_fn = tf.keras.losses.SparseCategoricalCrossentropy()
    compression = hvd.Compression.fp16 if a.fp16_allreduce else hvd.Compression.none

    def step(first):
        with tf.GradientTape() as tape:
            loss = loss_fn(labels, model(images, training=True))
        tape = hvd.DistributedGradientTape(tape, compression=compression)      # gradients are averaged across ranks here
        grads = tape.gradient(loss, model.trainable_variables)
        opt.apply_gradients(zip(grads, model.trainable_variables))
        if first:                                                               # after the first step the slots exist
            hvd.broadcast_variables(model.variables, root_rank=0)
            hvd.broadcast_variables(opt.variables() if callable(getattr(opt, 'variables', None)) else opt.variables, root_rank=0)

    for i in range(a.num_warmup_batches):
        step(i == 0)
    rates = []
    for _ in range(a.num_iters):
        t0 = time.perf_counter()
        for _ in range(a.num_batches_per_iter):
            step(False)
        rates.append(a.batch_size * a.num_batches_per_iter / (time.perf_counter() - t0))
        if hvd.rank() == 0:
            print('%.1f img/sec per rank' % rates[-1], flush=True)
    if hvd.rank() == 0:
        mean = sum(rates) / len(rates)
        print('Total img/sec on %d rank(s): %.1f' % (hvd.size(), hvd.size() * mean))
    hvd.shutdown()


if __name__ == '__main__':
    main()
