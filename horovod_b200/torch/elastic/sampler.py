"""ElasticSampler: a DistributedSampler that repartitions the *remaining* samples of an epoch when the world size
changes (API parity: horovod/torch/elastic/sampler.py)."""
import math
import random

import torch.utils.data.distributed

from horovod_b200.torch.mpi_ops import rank, size


class ElasticSampler(torch.utils.data.Sampler):
    """Usage: call `record_batch(batch_idx, batch_size)` after each processed batch, `set_epoch(epoch)` at the end of
    every epoch, and register the sampler in the TorchState."""

    def __init__(self, dataset, shuffle=True, seed=0):
        self.dataset = dataset
        self.shuffle = shuffle
        self.seed = seed
        self.epoch = 0
        self.processed_indices = set()
        self.num_replicas = 0
        self.rank = 0
        self.remaining_indices = []
        self.num_samples = 0
        self.total_size = 0
        self.reset()

    def set_epoch(self, epoch):
        """Sets the epoch: clears the processed indices and reshuffles."""
        self.epoch = epoch
        self.processed_indices = set()
        self.reset()

    def record_batch(self, batch_idx, batch_size):
        """Marks the samples of this rank's batch as processed so a later reset does not repeat them."""
        indices = set(self.get_indices(batch_idx, batch_size))
        self.processed_indices.update(indices)

    def get_indices(self, batch_idx, batch_size):
        start_idx = batch_idx * batch_size
        end_idx = min(start_idx + batch_size, len(self.indices))
        return self.indices[start_idx:end_idx]

    def load_state_dict(self, state_dict):
        self.epoch = state_dict['epoch']
        self.processed_indices = state_dict['processed_indices']
        self.reset()

    def state_dict(self):
        return dict(epoch=self.epoch, processed_indices=self.processed_indices)

    def reset(self):
        self.num_replicas = size()
        self.rank = rank()
        # exclude what the job already processed this epoch
        all_indices = [idx for idx in range(len(self.dataset)) if idx not in self.processed_indices]
        if self.shuffle:
            # shuffle deterministically from (seed, epoch) so every rank agrees
            random.Random(self.seed + self.epoch).shuffle(all_indices)
        self.remaining_indices = all_indices
        self.num_samples = int(math.ceil(len(self.remaining_indices) * 1.0 / self.num_replicas))
        self.total_size = self.num_samples * self.num_replicas

    def __iter__(self):
        self.indices = self.remaining_indices[:]
        # add extra samples to make it evenly divisible
        if self.indices:
            self.indices += self.indices[:(self.total_size - len(self.indices))]
        assert len(self.indices) == self.total_size
        # subsample
        self.indices = self.indices[self.rank:self.total_size:self.num_replicas]
        assert len(self.indices) == self.num_samples
        return iter(self.indices)

    def __len__(self):
        return self.num_samples
