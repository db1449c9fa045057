"""ElasticSampler: a distributed sampler that can be re-partitioned in the middle of an epoch.

Public API parity with horovod/torch/elastic/sampler.py (`set_epoch`, `record_batch`, `get_indices`, `state_dict`,
`load_state_dict`, `reset`, the attributes `processed_indices` / `remaining_indices` / `num_samples` / `total_size`).
After a reset (world size changed) only the samples nobody consumed yet are dealt out again, so an epoch still visits
every sample once (up to the padding that makes the shards equal).  The order is a permutation drawn with a
`torch.Generator` seeded by (seed, epoch): identical on every rank without communication.
"""
import torch
import torch.utils.data

from horovod_b200.torch.mpi_ops import rank, size


class ElasticSampler(torch.utils.data.Sampler):
    def __init__(self, dataset, shuffle=True, seed=0):
        self.dataset, self.shuffle, self.seed = dataset, shuffle, seed
        self.epoch = 0
        self.processed_indices = set()
        self.reset()

    # ---- bookkeeping the training loop drives ------------------------------------------------------------------------
    def set_epoch(self, epoch):
        """New epoch: nothing is consumed yet, new permutation."""
        self.epoch = epoch
        self.processed_indices = set()
        self.reset()

    def record_batch(self, batch_idx, batch_size):
        """Call after batch `batch_idx` (of this rank's iteration order) was processed."""
        self.processed_indices.update(self.get_indices(batch_idx, batch_size))

    def get_indices(self, batch_idx, batch_size):
        first = batch_idx * batch_size
        return self.indices[first:first + batch_size]

    def state_dict(self):
        return {'epoch': self.epoch, 'processed_indices': self.processed_indices}

    def load_state_dict(self, state_dict):
        self.epoch = state_dict['epoch']
        self.processed_indices = set(state_dict['processed_indices'])
        self.reset()

    # ---- partitioning ---------------------------------------------------------------------------------------------------
    def _epoch_order(self):
        n = len(self.dataset)
        if not self.shuffle:
            return list(range(n))
        g = torch.Generator()
        g.manual_seed(self.seed * 1_000_003 + self.epoch)
        return torch.randperm(n, generator=g).tolist()

    def reset(self):
        """(Re)computes this rank's share of the not-yet-consumed samples for the current world."""
        self.num_replicas, self.rank = size(), rank()
        done = self.processed_indices
        self.remaining_indices = [i for i in self._epoch_order() if i not in done]
        self.num_samples = -(-len(self.remaining_indices) // self.num_replicas)
        self.total_size = self.num_samples * self.num_replicas
        padded = list(self.remaining_indices)
        while padded and len(padded) < self.total_size:  # wrap around so every rank gets num_samples indices
            padded += padded[:self.total_size - len(padded)]
        self.indices = padded[self.rank:self.total_size:self.num_replicas]

    def __iter__(self):
        return iter(list(self.indices))

    def __len__(self):
        return self.num_samples
