"""Keras `model.fit` with `hvd.DistributedOptimizer` and the hvd callbacks (synthetic MNIST-shaped data).

    hvdrun -np 4 python examples/tensorflow2_keras_synthetic.py --epochs 3

Needs TensorFlow >= 2.4 (see tensorflow2_synthetic_benchmark.py about the build image).
"""
import argparse

import numpy as np
import tensorflow as tf

import horovod_b200.tensorflow.keras as hvd


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--epochs', type=int, default=3)
    p.add_argument('--batch-size', type=int, default=128)
    p.add_argument('--checkpoint', default=None, help='rank 0 writes the best model here')
    a = p.parse_args()

    hvd.init()
    gpus = tf.config.list_physical_devices('GPU')
    if gpus:
        tf.config.set_visible_devices(gpus[hvd.local_rank()], 'GPU')

    rng = np.random.RandomState(hvd.rank())                     # every rank draws its own shard
    x = rng.rand(4096, 28, 28, 1).astype('float32')
    y = (x.mean(axis=(1, 2, 3)) * 20).astype('int64') % 10
    model = tf.keras.Sequential([
        tf.keras.layers.Conv2D(16, 3, activation='relu', input_shape=(28, 28, 1)),
        tf.keras.layers.MaxPooling2D(),
        tf.keras.layers.Flatten(),
        tf.keras.layers.Dense(64, activation='relu'),
        tf.keras.layers.Dense(10, activation='softmax')])
    # scale the learning rate by the number of ranks, then warm it up from the single-rank value
    base_lr = 0.001
    opt = hvd.DistributedOptimizer(tf.keras.optimizers.Adam(base_lr * hvd.size()))
    model.compile(optimizer=opt, loss='sparse_categorical_crossentropy', metrics=['accuracy'])
    callbacks = [
        hvd.callbacks.BroadcastGlobalVariablesCallback(0),       # same initial state everywhere
        hvd.callbacks.MetricAverageCallback(),                   # epoch metrics averaged over ranks
        hvd.callbacks.LearningRateWarmupCallback(initial_lr=base_lr * hvd.size(), warmup_epochs=1, verbose=0),
    ]
    if a.checkpoint and hvd.rank() == 0:
        callbacks.append(hvd.callbacks.BestModelCheckpoint(monitor='loss', filepath=a.checkpoint))
    model.fit(x, y, batch_size=a.batch_size, epochs=a.epochs, callbacks=callbacks, verbose=1 if hvd.rank() == 0 else 0)
    hvd.shutdown()


if __name__ == '__main__':
    main()
