"""This rank's share of a Parquet dataset, decoded to numpy columns.

The reference reads its intermediate Parquet through Petastorm readers (`make_batch_reader(cur_shard=rank,
shard_count=size)`, horovod/spark/torch/remote.py:240-300).  Here the files of the dataset are dealt to ranks
round-robin (`shard_files`), a rank's files are decoded with pyarrow into ONE dict of contiguous numpy arrays (list
columns become [rows, ...] arrays, reshaped by the declared per-row shapes), and every rank derives the SAME number of
steps per epoch from the smallest shard so that no rank runs out of collectives early.
"""
import numpy as np


def shard_files(paths, rank, size):
    """Round-robin deal; a rank that would get nothing (more ranks than files) shares a file with another rank."""
    paths = sorted(paths)
    mine = [p for i, p in enumerate(paths) if i % size == rank]
    return mine or [paths[rank % len(paths)]]


def _column_to_numpy(chunked, row_shape=None):
    import pyarrow as pa
    typ = chunked.type
    if pa.types.is_list(typ) or pa.types.is_large_list(typ) or pa.types.is_fixed_size_list(typ):
        rows = chunked.to_pylist()
        try:
            arr = np.asarray(rows)
        except ValueError:
            arr = np.empty(len(rows), dtype=object)
            arr[:] = rows
        if arr.dtype == np.float64:
            arr = arr.astype(np.float32)
    else:
        arr = chunked.to_numpy(zero_copy_only=False) if hasattr(chunked, 'to_numpy') else np.asarray(chunked.to_pylist())
        if arr.dtype == np.float64:
            arr = arr.astype(np.float32)
    if row_shape is not None and arr.dtype != object:
        arr = arr.reshape([len(arr)] + [d for d in row_shape if d != -1])
    return arr


class ParquetShard:
    """rows of rank `rank` out of `size`.  `row_shapes`: {column: per-row shape}; `steps(batch_size)` is identical on
    every rank."""

    def __init__(self, store, path, columns, rank=0, size=1, row_shapes=None):
        import pyarrow.dataset as ds
        self.columns = list(columns)
        self.row_shapes = dict(row_shapes or {})
        dataset = ds.dataset(store._local(path), format='parquet', filesystem=store.fs)
        frags = sorted(dataset.get_fragments(), key=lambda f: f.path)
        if not frags:
            raise ValueError('no Parquet files under %s' % path)
        by_path = {f.path: f for f in frags}
        self._frags = [by_path[p] for p in shard_files(list(by_path), rank, size)]
        counts = {f.path: f.count_rows() for f in frags}
        per_rank = [sum(counts[p] for p in shard_files(list(by_path), r, size)) for r in range(size)]
        self.rows = per_rank[rank]
        self.min_rows_per_rank = min(per_rank)
        self.total_rows = sum(counts.values())
        self._data = None

    def steps(self, batch_size):
        return max(1, self.min_rows_per_rank // batch_size)

    def load(self):
        """{column: numpy array of this rank's rows}; decoded once, then cached."""
        if self._data is None:
            import pyarrow as pa
            table = pa.concat_tables([f.to_table(columns=self.columns) for f in self._frags])
            self._data = {c: _column_to_numpy(table.column(c), self.row_shapes.get(c)) for c in self.columns}
        return self._data

    def release(self):
        self._data = None
